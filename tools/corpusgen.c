/*
 * corpusgen.c -- deterministic synthetic-document generator for bench.py and the tests
 * (SURVEY 8d).  Test/bench tooling, not product code.
 *
 *   RNG       splitmix64, state0 = seed
 *   make_doc  append pool[rng % n_lines] joined by a single ' ' until length >= L, then cut to
 *             the largest prefix <= L bytes that ends on a code-point boundary (a split
 *             code point would make the whole document invalid UTF-8 -> 0 ids)
 *   L         fixed (cfg 2/4/5) or floor(64 * 64^u), u uniform [0,1) (cfg 3: log-uniform 64..4096)
 *   emoji     when emoji_every > 0, every emoji_every-th document gets U+1F600 + (rng & 0x3F)
 *             inserted after its first line (cfg 4: exercises 4-byte code points)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* Returns total bytes written, or -(needed capacity estimate) if out_cap is too small. */
int64_t gen_docs(const uint8_t* pool, const int64_t* line_off, int64_t n_lines, uint64_t seed,
                 int64_t n_docs, int64_t fixed_len, int emoji_every,
                 uint8_t* out, int64_t out_cap, int64_t* doc_off) {
    uint64_t st = seed;
    int64_t w = 0;
    doc_off[0] = 0;
    for (int64_t d = 0; d < n_docs; ++d) {
        int64_t L = fixed_len;
        if (fixed_len <= 0) {
            const double u = (double)(splitmix64(&st) >> 11) * (1.0 / 9007199254740992.0);
            L = (int64_t)floor(64.0 * pow(64.0, u));
        }
        if (w + L + 8 > out_cap) return -(w + L + 8);
        const int64_t start = w;
        int first = 1;
        while (w - start < L) {
            const int64_t li = (int64_t)(splitmix64(&st) % (uint64_t)n_lines);
            const int64_t a = line_off[li], b = line_off[li + 1];
            if (!first) out[w++] = ' ';
            int64_t room = (start + L + 8) - w;           /* a little slack; cut below */
            int64_t n = b - a; if (n > room) n = room;
            if (n > 0) { memcpy(out + w, pool + a, (size_t)n); w += n; }
            if (first && emoji_every > 0 && (d % emoji_every) == 0 && w + 4 <= start + L + 8) {
                const uint32_t cp = 0x1F600u + (uint32_t)(splitmix64(&st) & 0x3F);
                out[w++] = (uint8_t)(0xF0 | (cp >> 18)); out[w++] = (uint8_t)(0x80 | ((cp >> 12) & 0x3F));
                out[w++] = (uint8_t)(0x80 | ((cp >> 6) & 0x3F)); out[w++] = (uint8_t)(0x80 | (cp & 0x3F));
            }
            first = 0;
            if (n_lines <= 0) break;
        }
        /* cut to <= L bytes on a code-point boundary */
        int64_t end = start + L; if (end > w) end = w;
        while (end > start && end < w && (out[end] & 0xC0) == 0x80) --end;
        w = end;
        doc_off[d + 1] = w;
    }
    return w;
}
