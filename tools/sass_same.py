#!/usr/bin/env python3
"""Is a kernel's machine code the same in two builds of the library?  Disassembles the kernel from both shared objects
(cuobjdump -xelf, nvdisasm), renumbers the branch labels in order of appearance, ignores white space, and prints the
lines that differ.  Used when a change is meant to leave a benched kernel alone (a new template instantiation next to it):

  python tools/sass_same.py /tmp/lib_before.so blingfire_b200/lib/libblingfiretokdll.so sp_unigram_kernelE sp_bpe_kernelE
"""
import os
import re
import subprocess
import sys
import tempfile


def kernel_sass(lib, name):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        out = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines()
        res, inside = [], False
        for ln in out:
            if ln.startswith("//---") and ".text." in ln:
                inside = name in ln
            if inside:
                res.append(re.sub(r"/\*[0-9a-f]+\*/", "", ln))
        if res:
            labels = {}
            res = [re.sub(r"\.L_x_\d+", lambda m: labels.setdefault(m.group(0), f".L{len(labels)}"), l) for l in res]
            res = [re.sub(r"\$__internal_\d+_", "$__internal_N_", l) for l in res]      # the compiler's numbering of its helpers
            return [re.sub(r"\s+", " ", l).strip() for l in res]
    raise SystemExit(f"{name}: not found in {lib}")


def main():
    a, b, names = sys.argv[1], sys.argv[2], sys.argv[3:]
    rc = 0
    for n in names:
        x, y = kernel_sass(a, n), kernel_sass(b, n)
        diff = [(i, p, q) for i, (p, q) in enumerate(zip(x, y)) if p != q]
        same_multiset = sorted(x) == sorted(y)
        print(f"{n}: {len(x)} / {len(y)} lines, {len(diff)} differ" + (" (the same instructions, reordered)" if diff and same_multiset else ""))
        for i, p, q in diff[:8]:
            print(f"  {i}: < {p[:100]}\n  {i}: > {q[:100]}")
        if len(x) != len(y) or (diff and not same_multiset):
            rc = 1
    sys.exit(rc)


if __name__ == "__main__":
    main()
