#!/usr/bin/env python3
"""Per-source-line view of an ncu capture: joins the SASS page of `ncu --page source --csv` (instruction
counts and stall samples per instruction) with the line table nvdisasm prints for the same cubin.

  python tools/ncu_lines.py gpurun_out/r02b_wp_tokenize_source.csv blingfire_b200/lib/libblingfiretokdll.so wp_tokenize_kernelItE [top]

The binary must be the one that was profiled (same build).  Output: one row per file:line, sorted by warp
instructions executed, with the share of the kernel's instructions and of its stall samples.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_table(lib, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kernel not in out:
            continue
        table = {}
        inside = False
        cur = ("?", 0)
        for ln in out.splitlines():
            if ln.startswith("//---") and ".text." in ln:
                inside = kernel in ln
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                table[int(m.group(1), 16)] = (cur, m.group(2).strip())
        if table:
            return table
    raise SystemExit(f"kernel {kernel} not found in {lib}")


def main():
    src_csv, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(lib, kernel)
    rows = list(csv.reader(open(src_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    base = None
    per_line = defaultdict(lambda: [0, 0, 0, defaultdict(int)])
    tot_inst = tot_samp = 0
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        if base is None:
            base = a
        off = a - base
        (fl, sass) = table.get(off, (("?", 0), r[col["Source"]]))
        inst = int(float(r[col["Instructions Executed"]] or 0))
        samp = int(float(r[col["# Samples"]] or 0))
        thr = int(float(r[col["Thread Instructions Executed"]] or 0))
        e = per_line[fl]
        e[0] += inst; e[1] += samp; e[2] += thr
        for n in stall_cols:
            v = r[col[n]]
            if v and v != "0":
                e[3][n] += int(float(v))
        tot_inst += inst; tot_samp += samp
    print(f"kernel {kernel}: {tot_inst} warp instructions, {tot_samp} samples, {len(table)} SASS instructions")
    print(f"{'file:line':28s} {'warp-inst':>12s} {'%inst':>6s} {'thr/inst':>8s} {'samples':>8s} {'%samp':>6s}  top stalls")
    for fl, e in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        stalls = ", ".join(f"{k[6:]}={v}" for k, v in sorted(e[3].items(), key=lambda kv: -kv[1])[:3])
        print(f"{fl[0] + ':' + str(fl[1]):28s} {e[0]:12d} {100 * e[0] / max(tot_inst, 1):6.2f} {e[2] / max(e[0], 1):8.1f} {e[1]:8d} "
              f"{100 * e[1] / max(tot_samp, 1):6.2f}  {stalls}")


if __name__ == "__main__":
    main()
