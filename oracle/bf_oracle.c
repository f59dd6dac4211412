/*
 * bf_oracle.c -- TEST INFRASTRUCTURE ONLY (see bf_oracle.h).
 *
 * Plain-C restatement of the reference's CPU algorithm for the TextToIds /
 * TextToWords hot path.  Every function cites the reference file:line it
 * follows (paths relative to the reference checkout).  It works on the packed
 * .bin image exactly like the reference readers do (state == byte offset into
 * the automaton dump, variable-length records), which keeps it independent of
 * the product's flattened HBM tables.
 *
 * Parity: pinned against the reference's documented known-answer vectors and
 * against the reference library built from its own sources (oracle/Makefile ->
 * oracle/_ref/libblingfiretokdll.so); see tests/test_oracle.py.
 */
#include "bf_oracle.h"

#include <float.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants: blingfireclient.library/inc/FAFsmConst.h:68-75,152-273,365-371 ---- */
enum {
    IW_ANY = 0, IW_L_ANCHOR = 1, IW_R_ANCHOR = 2, IW_EPSILON = 3,
    DFA_DEAD_STATE = -2,
    TRS_NONE = 0, TRS_RANGE = 1, TRS_IMPL = 2, TRS_PARA = 4, TRS_IWIA = 6,
    FUNC_POS_DICT = 12, FUNC_WBD = 19, FUNC_GLOBAL = 20, FUNC_I2W = 35,
    PARAM_FSM = 2, PARAM_REVERSE = 10, PARAM_DIRECTION = 11, PARAM_MAP_MODE = 16,
    PARAM_NO_TR = 18, PARAM_IGNORE_CASE = 22, PARAM_ARRAY = 24, PARAM_MULTI_MAP = 25,
    PARAM_FSM_TYPE = 26, PARAM_DICT_MODE = 31, PARAM_NORMALIZE = 35, PARAM_DO_W2B = 37,
    PARAM_DEPTH = 38, PARAM_MAX_TAG = 39, PARAM_LOG_SCALE = 40, PARAM_WORD = 42,
    PARAM_PUNKT = 43, PARAM_EOS = 44, PARAM_EOP = 45, PARAM_USE_NFST = 46,
    PARAM_CHARMAP = 47, PARAM_XWORD = 51, PARAM_SEG = 52, PARAM_IGNORE = 53,
    PARAM_ACT_DATA = 68, PARAM_MAX_LENGTH = 69, PARAM_VERIFY_LDB_BIN = 70,
    PARAM_TOKENIZATION_TYPE = 71, PARAM_ID_OFFSET = 72, PARAM_USE_BYTE_ENCODING = 73,
    PARAM_NO_DUMMY_PREFIX = 74, PARAM_STRING_ARRAY = 75, PARAM_TOKENID_MIN = 76, PARAM_TOKENID_MAX = 77,
    TYPE_MOORE_DFA = 3, TYPE_MEALY_DFA = 7,
    MODE_PACK_TRIV = 1, MODE_PACK_MPH = 2, MODE_PACK_FIXED = 3,
    TOKENIZE_BPE = 3, TOKENIZE_BPE_OPT = 4, TOKENIZE_BPE_OPT_WITH_MERGES = 5,
    MAX_ARR_SIZE = 1000000000,  /* FALimits.h:25 */
    MAX_WORD_LEN = 300,         /* FALimits.h:35 */
    DEF_MAX_DEPTH = 2,          /* FALexTools_t.h:108 */
    WBD_WORD_TAG = 1, WBD_IGNORE_TAG = 4, /* blingfiretokdll.cpp:38-39 */
    SP_DELIM = 0x2581           /* blingfiretokdll.h:11 */
};

typedef unsigned char u8;

/* ---- packed containers ---- */

/* FAChains_pack_triv.h: {int SizeOfValue; int MaxCount; chains...} */
typedef struct { const u8* img; int size_of_value; int max_count; } chains_t;

/* FAMultiMap_pack.cpp:22-53 */
typedef struct { const u8* offsets; unsigned max_key; int size_of_offset; chains_t values; int valid; } mmap_t;

/* FAMultiMap_pack_fixed.cpp:25-58 */
typedef struct { const u8* data; int size_of_value, size_of_arr, max_count, min_key, max_key; int valid; } mmap_fixed_t;

/* FAIwMap_pack.cpp:35-62 */
typedef struct { int interval_count; const int* from_iw; const int* to_iw_offset; int size_of_new_iw; const u8* new_iws; } iwmap_t;

/* FARSDfa_pack_triv.cpp:27-77 (+ FAState2Ow_pack_triv, FAMealyDfa_pack_triv share the image) */
typedef struct {
    const u8* img; int dst_size; int initial; int remap; iwmap_t iwmap;
    chains_t ows; int has_ows;
} dfa_t;

struct bfo_model {
    u8* image; long image_size;
    int dump_count; const u8* dumps[256]; int dump_off[256];
    mmap_t conf;
    /* [wbd] */
    int has_wbd; dfa_t wbd_dfa; mmap_t acts; int has_acts; mmap_fixed_t wbd_charmap; int has_wbd_charmap;
    int max_depth, max_token_length, ignore_case;
    int* fn2ini; int fn2ini_size;
    /* [pos-dict] */
    int has_seg; dfa_t seg_dfa; mmap_fixed_t i2info_fixed; mmap_t i2info_triv; int i2info_mode;
    mmap_fixed_t seg_charmap; int has_seg_charmap;
    int tok_algo, id_offset, use_raw_bytes, no_dummy_prefix;
    /* [i2w] (blingfiretokdll.cpp:997-1045): FAStringArray_pack image, regular id range */
    int has_i2w; int i2w_count; const u8* i2w_offsets; const u8* i2w_data; int min_token_id, max_token_id;
};

static int rd_i32(const u8* p) { int v; memcpy(&v, p, 4); return v; }
static unsigned rd_u32(const u8* p) { unsigned v; memcpy(&v, p, 4); return v; }

/* FAEncodeUtils.h:292-310 (FADecode_UC_US_UI): native little-endian */
static unsigned dec_uc_us_ui(const u8* p, int size) {
    if (size == 1) return p[0];
    if (size == 2) { unsigned short v; memcpy(&v, p, 2); return v; }
    return rd_u32(p);
}
/* FAEncodeUtils.h:418-448 (FADecode_1_2_3_4_idx): big-endian */
static unsigned dec_be_idx(const u8* p, unsigned idx, int size) {
    const u8* q = p + (size_t)idx * size; unsigned v = 0;
    for (int i = 0; i < size; ++i) v = (v << 8) | q[i];
    return v;
}
/* FAEncodeUtils.h:456-503 (FADecodeDst_idx): big-endian, all-ones = dead state */
static int dec_dst_idx(const u8* p, unsigned idx, int size) {
    unsigned v = dec_be_idx(p, idx, size);
    unsigned ones = (size == 4) ? 0xffffffffu : ((1u << (8 * size)) - 1u);
    if (v == ones) return DFA_DEAD_STATE;
    return (int)v;
}

static void chains_set(chains_t* c, const u8* img) {  /* FAChains_pack_triv.cpp:21-30 */
    c->img = img; c->size_of_value = rd_i32(img); c->max_count = rd_i32(img + 4);
}
/* FAChains_pack_triv.h:144-163 (UnPack by pointer: only int-sized values) */
static int chains_unpack_ptr(const chains_t* c, int off, const int** vals) {
    if (c->size_of_value != 4) return -1;
    *vals = (const int*)(c->img + off + 4);
    return rd_i32(c->img + off);
}
/* FAChains_pack_triv.h:166-222 (UnPack by index) */
static int chains_unpack_idx(const chains_t* c, int off, int idx) {
    const u8* p = c->img + off;
    if (c->size_of_value == 1) {
        int cnt = (signed char)p[0];
        if (idx < cnt) return (signed char)p[1 + idx];
    } else if (c->size_of_value == 2) {
        short cnt; memcpy(&cnt, p, 2);
        if (idx < cnt) { short v; memcpy(&v, p + 2 + 2 * idx, 2); return v; }
    } else {
        int cnt = rd_i32(p);
        if (idx < cnt) return rd_i32(p + 4 + 4 * idx);
    }
    return -1;
}

static void mmap_set(mmap_t* m, const u8* d) {  /* FAMultiMap_pack.cpp:22-53 */
    unsigned off = 0;
    m->max_key = rd_u32(d); off += 4;
    m->size_of_offset = (int)rd_u32(d + off); off += 4;
    m->offsets = d + off; off += m->size_of_offset * (1 + m->max_key);
    if (off % 4) off += 4 - off % 4;
    chains_set(&m->values, d + off);
    m->valid = 1;
}
/* FAMultiMap_pack.cpp:106-126 (Get by pointer) */
static int mmap_get(const mmap_t* m, int key, const int** vals) {
    if (key < 0 || m->max_key < (unsigned)key) return -1;
    unsigned vo = dec_be_idx(m->offsets, (unsigned)key, m->size_of_offset);
    if (vo == 0) return -1;
    return chains_unpack_ptr(&m->values, (int)(vo - 1), vals);
}

static void mmapf_set(mmap_fixed_t* m, const u8* d) {  /* FAMultiMap_pack_fixed.cpp:25-58 */
    m->size_of_value = (int)rd_u32(d);
    m->max_count = rd_i32(d + 4);
    m->size_of_arr = (m->max_count + 1) * m->size_of_value;
    m->min_key = rd_i32(d + 8);
    m->max_key = rd_i32(d + 12);
    m->data = d + 16;
    m->valid = 1;
}
/* FAMultiMap_pack_fixed.cpp:67-137 (Get, copying) */
static int mmapf_get_copy(const mmap_fixed_t* m, int key, int* out, int max_out) {
    if (key < m->min_key || key > m->max_key) return -1;
    const u8* arr = m->data + (size_t)m->size_of_arr * (unsigned)(key - m->min_key);
    int cnt;
    if (m->size_of_value == 1) {
        cnt = (signed char)arr[0];
        if (cnt > m->max_count) return -1;
        if (out && max_out >= cnt) for (int i = 0; i < cnt; ++i) out[i] = (signed char)arr[1 + i];
    } else if (m->size_of_value == 2) {
        short c; memcpy(&c, arr, 2); cnt = c;
        if (cnt > m->max_count) return -1;
        if (out && max_out >= cnt) for (int i = 0; i < cnt; ++i) { short v; memcpy(&v, arr + 2 + 2 * i, 2); out[i] = v; }
    } else {
        cnt = rd_i32(arr);
        if (cnt > m->max_count) return -1;
        if (out && max_out >= cnt && cnt > 0) memcpy(out, arr + 4, sizeof(int) * (size_t)cnt);
    }
    return cnt;
}
/* FAMultiMap_pack_fixed.cpp:140-162 (Get by pointer: only int-sized values) */
static int mmapf_get_ptr(const mmap_fixed_t* m, int key, const int** vals) {
    if (key < m->min_key || key > m->max_key || m->size_of_value != 4) return -1;
    const u8* arr = m->data + (size_t)m->size_of_arr * (unsigned)(key - m->min_key);
    int cnt = rd_i32(arr);
    if (cnt > m->max_count) return -1;
    *vals = (const int*)(arr + 4);
    return cnt;
}

static void iwmap_set(iwmap_t* w, const u8* img) {  /* FAIwMap_pack.cpp:35-62 */
    unsigned off = 0;
    w->size_of_new_iw = rd_i32(img); off += 4;
    w->interval_count = rd_i32(img + off); off += 4;
    w->from_iw = (const int*)(img + off); off += 4 * w->interval_count;
    w->to_iw_offset = (const int*)(img + off); off += 8 * w->interval_count;
    w->new_iws = img + off;
}
/* FAIwMap_pack.h:55-109 (GetNewIw; the int cache is only a speed-up) */
static int iwmap_get(const iwmap_t* w, int old_iw) {
    int idx = -1;  /* FAFindEqualOrLess_log: last interval with From <= OldIw */
    int lo = 0, hi = w->interval_count - 1;
    while (lo <= hi) { int mid = (lo + hi) >> 1; if (w->from_iw[mid] <= old_iw) { idx = mid; lo = mid + 1; } else hi = mid - 1; }
    if (idx == -1) return -1;
    int from = w->from_iw[idx], end = w->to_iw_offset[2 * idx], ioff = w->to_iw_offset[2 * idx + 1];
    if (old_iw > end) return -1;
    unsigned v = dec_be_idx(w->new_iws + ioff, (unsigned)(old_iw - from), w->size_of_new_iw);
    return v ? (int)v - 1 : -1;
}

/* FARSDfa_pack_triv.cpp:27-77 / FAMealyDfa_pack_triv.cpp:24-67 (SetImage) */
static int dfa_set(dfa_t* d, const u8* img, int mealy) {
    unsigned off = 0;
    d->img = img;
    d->dst_size = rd_i32(img); off += 4;
    if (d->dst_size < 1 || d->dst_size > 4) d->dst_size = 3;
    int ows_off = rd_i32(img + off); off += 4;
    unsigned iwc = rd_u32(img + off); off += 4;
    d->remap = (iwc & 0x80000000u) != 0; iwc &= 0x7fffffffu;
    off += 4 * iwc;
    d->has_ows = 0;
    if (mealy) {
        if (ows_off == 0 || d->remap) return 0;
        chains_set(&d->ows, img + ows_off); d->has_ows = 1;
    } else if (d->remap) {
        int sz = rd_i32(img + off); off += 4;
        iwmap_set(&d->iwmap, img + off); off += sz;
    }
    d->initial = (int)off;
    return 1;
}
static int dfa_is_final(const dfa_t* d, int s) {  /* FARSDfa_pack_triv.cpp:128-138 */
    if (s < 0) return 0;
    return (d->img[s] & 0x80) != 0;
}
/* exact-match search in a sorted little-endian Iw array (FAFind_log, FAUtils_cl.h:86-140) */
static int find_iw(const u8* iws, unsigned n, int iw_size, unsigned val) {
    int lo = 0, hi = (int)n - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1; unsigned cur = dec_uc_us_ui(iws + (size_t)mid * iw_size, iw_size);
        if (cur == val) return mid;
        if (val < cur) hi = mid - 1; else lo = mid + 1;
    }
    return -1;
}
/* last index with from[idx] <= val (FAFindEqualOrLess_log, FAUtils_cl.h:143-200) */
static int find_le(const u8* iws, unsigned n, int iw_size, unsigned val) {
    int lo = 0, hi = (int)n - 1, idx = -1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1; unsigned cur = dec_uc_us_ui(iws + (size_t)mid * iw_size, iw_size);
        if (cur <= val) { idx = mid; lo = mid + 1; } else hi = mid - 1;
    }
    return idx;
}
/* FARSDfa_pack_triv.cpp:141-399 (GetDest) */
static int dfa_get_dest(const dfa_t* d, int state, int iw) {
    if (state < 0) return -1;
    int niw = iw;
    if (d->remap) { niw = iwmap_get(&d->iwmap, iw); if (niw == -1) return -1; }
    const u8* p = d->img + state;
    const u8 info = *p++;
    const int iw_size = ((info & 0x18) >> 3) + 1;
    const int tr = info & 0x07;
    const unsigned mask = iw_size == 1 ? 0xFFFFFF00u : (iw_size == 2 ? 0xFFFF0000u : 0u);
    switch (tr) {
    case TRS_PARA: {
        if (mask & (unsigned)niw) return -1;
        unsigned cnt = 1 + dec_uc_us_ui(p, iw_size); p += iw_size;
        int idx = find_iw(p, cnt, iw_size, (unsigned)niw);
        p += (size_t)cnt * iw_size;
        if (idx == -1) return -1;
        return dec_dst_idx(p, (unsigned)idx, d->dst_size);
    }
    case TRS_IWIA: {
        unsigned base = dec_uc_us_ui(p, iw_size); p += iw_size;
        unsigned mx = dec_uc_us_ui(p, iw_size); p += iw_size;
        if (niw < (int)base || niw > (int)mx) return -1;
        int dst = dec_dst_idx(p, (unsigned)(niw - (int)base), d->dst_size);
        return dst == 0 ? -1 : dst;
    }
    case TRS_RANGE: {
        if (mask & (unsigned)niw) return -1;
        unsigned cnt = 1 + dec_uc_us_ui(p, iw_size); p += iw_size;
        int idx = find_le(p, cnt, iw_size, (unsigned)niw);
        if (idx == -1) return -1;
        p += (size_t)cnt * iw_size;
        if (dec_uc_us_ui(p + (size_t)idx * iw_size, iw_size) < (unsigned)niw) return -1;
        p += (size_t)cnt * iw_size;
        return dec_dst_idx(p, (unsigned)idx, d->dst_size);
    }
    case TRS_IMPL: {
        int owc = (info & 0x60) >> 5; int ow_size = owc == 3 ? 4 : owc;
        if ((unsigned)niw == dec_uc_us_ui(p, iw_size)) return state + 1 + iw_size + ow_size;
        return -1;
    }
    default: return -1;
    }
}
/* FAState2Ow_pack_triv.cpp:34-130 (GetOw) */
static int dfa_get_ow(const dfa_t* d, int state) {
    const u8* p = d->img + state;
    const u8 info = *p++;
    const int owc = (info & 0x60) >> 5;
    if (owc == 0) return -1;
    const int iw_size = ((info & 0x18) >> 3) + 1;
    switch (info & 0x07) {
    case TRS_PARA: { unsigned c = dec_uc_us_ui(p, iw_size); p += iw_size; p += (size_t)(c + 1) * (d->dst_size + iw_size); break; }
    case TRS_IWIA: { unsigned b = dec_uc_us_ui(p, iw_size); p += iw_size; unsigned m = dec_uc_us_ui(p, iw_size); p += iw_size;
                     p += (size_t)d->dst_size * (m - b + 1); break; }
    case TRS_RANGE: { unsigned c = dec_uc_us_ui(p, iw_size); p += iw_size; p += (size_t)(c + 1) * (d->dst_size + 2 * iw_size); break; }
    case TRS_IMPL: p += iw_size; break;
    default: break;
    }
    if (owc == 1) return (signed char)p[0];
    if (owc == 2) { short v; memcpy(&v, p, 2); return v; }
    return rd_i32(p);
}
/* FAMealyDfa_pack_triv.cpp:69-244 (GetDestOw; only PARA and IMPL are supported there) */
static int mealy_get_dest_ow(const dfa_t* d, int state, int iw, int* ow) {
    if (state < 0) return -1;
    const u8* p = d->img + state;
    const u8 info = *p++;
    const int iw_size = ((info & 0x18) >> 3) + 1;
    const int owc = (info & 0x60) >> 5;
    const unsigned mask = iw_size == 1 ? 0xFFFFFF00u : (iw_size == 2 ? 0xFFFF0000u : 0u);
    const u8* ows_ptr = NULL; int idx, dst;
    switch (info & 0x07) {
    case TRS_PARA: {
        if (mask & (unsigned)iw) return -1;
        unsigned cnt = 1 + dec_uc_us_ui(p, iw_size); p += iw_size;
        idx = find_iw(p, cnt, iw_size, (unsigned)iw);
        p += (size_t)cnt * iw_size;
        if (idx == -1) return -1;
        if (owc != 0) ows_ptr = p + (size_t)d->dst_size * cnt;
        dst = dec_dst_idx(p, (unsigned)idx, d->dst_size);
        break;
    }
    case TRS_IMPL: {
        idx = 0;
        int ow_size = owc == 3 ? 4 : owc;
        if ((unsigned)iw != dec_uc_us_ui(p, iw_size)) return -1;
        ows_ptr = p + iw_size;
        dst = state + 1 + iw_size + ow_size;
        break;
    }
    default: return -1;
    }
    if (owc > 0 && ows_ptr) {
        int off;
        if (owc == 1) off = (signed char)ows_ptr[0];
        else if (owc == 2) { short v; memcpy(&v, ows_ptr, 2); off = v; }
        else off = rd_i32(ows_ptr);
        *ow = chains_unpack_idx(&d->ows, off, idx);
    } else {
        *ow = -1;
    }
    return dst;
}

/* FAUtils_cl.cpp:148-159 (FAGetCrc32): standard reflected CRC-32, chained across dumps */
static unsigned crc32_update(const u8* buf, size_t size, unsigned crc) {
    static unsigned table[256]; static int init = 0;
    if (!init) {
        for (unsigned i = 0; i < 256; ++i) { unsigned c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
        init = 1;
    }
    crc ^= ~0u;
    while (size--) crc = table[(crc ^ *buf++) & 0xFF] ^ (crc >> 8);
    return crc ^ ~0u;
}

static int is_boolean_param(int p) {  /* FALDB.cpp:130-141 */
    return p == PARAM_REVERSE || p == PARAM_NO_TR || p == PARAM_IGNORE_CASE || p == PARAM_DICT_MODE ||
           p == PARAM_NORMALIZE || p == PARAM_LOG_SCALE || p == PARAM_USE_NFST || p == PARAM_DO_W2B ||
           p == PARAM_VERIFY_LDB_BIN;
}
/* FALDB.cpp:143-190 (GetValue) */
static int ldb_get_value(const bfo_model* m, int section, int param, int* value) {
    *value = 0;
    const int* vals = NULL; int size = mmap_get(&m->conf, section, &vals);
    for (int i = 0; i < size; ++i) {
        int np = vals[i]; int isb = is_boolean_param(np);
        if (!isb) { i++; if (i >= size) return 0; }
        if (np == param) { *value = isb ? 1 : vals[i]; return 1; }
    }
    return is_boolean_param(param);
}
/* FALDB.cpp:67-116 (IsValidBinary) */
static int ldb_is_valid(const bfo_model* m) {
    int verify = 0; ldb_get_value(m, FUNC_GLOBAL, PARAM_VERIFY_LDB_BIN, &verify);
    if (!verify) return 1;
    if (m->dump_count < 2) return 0;
    const u8* v = m->dumps[m->dump_count - 1];
    if (rd_u32(v) != 0) return 1;
    unsigned exp_size = rd_u32(v + 4), exp_hash = rd_u32(v + 8), size = 0, hash = 0;
    for (int i = 0; i < m->dump_count - 1; ++i) {
        int s = m->dump_off[i + 1] - m->dump_off[i];
        if (s < 0) return 0;
        size += (unsigned)s; hash = crc32_update(m->dumps[i], (size_t)s, hash);
    }
    return size == exp_size && hash == exp_hash;
}

/* FAWbdConfKeeper.cpp:246-314 (CalcFnIniStates) */
static void calc_fn_ini(bfo_model* m) {
    const int initial = m->wbd_dfa.initial;
    const int state_r = dfa_get_dest(&m->wbd_dfa, initial, IW_R_ANCHOR);
    if (state_r == -1) return;
    const int* act; int n, id = 0, max_fn = -1;
    while ((n = mmap_get(&m->acts, id++, &act)) != -1) {
        int i = 2;
        for (; i < n; ++i) if (act[i] == 0 && i + 1 < n) { i++; break; }
        for (; i < n; ++i) if (max_fn < act[i]) max_fn = act[i];
    }
    if (max_fn == -1) return;
    m->fn2ini_size = max_fn + 1;
    m->fn2ini = (int*)malloc(sizeof(int) * (size_t)m->fn2ini_size);
    m->fn2ini[0] = initial;
    for (int f = 1; f <= max_fn; ++f) m->fn2ini[f] = dfa_get_dest(&m->wbd_dfa, state_r, f);
}

/* FAWbdConfKeeper.cpp:56-232 (Initialize) */
static int init_wbd(bfo_model* m, const int* v, int n) {
    const u8* fsm = NULL;
    m->max_depth = DEF_MAX_DEPTH; m->max_token_length = MAX_WORD_LEN; m->ignore_case = 0;
    for (int i = 0; i < n; ++i) {
        switch (v[i]) {
        case PARAM_MAP_MODE: if (v[++i] != MODE_PACK_TRIV) return 0; break;
        case PARAM_DEPTH: m->max_depth = v[++i]; break;
        case PARAM_MAX_LENGTH: m->max_token_length = v[++i]; break;
        case PARAM_IGNORE_CASE: m->ignore_case = 1; break;
        case PARAM_FSM_TYPE: if (v[++i] != TYPE_MOORE_DFA) return 0; break;
        case PARAM_FSM: fsm = m->dumps[v[++i]]; if (!dfa_set(&m->wbd_dfa, fsm, 0)) return 0; break;
        case PARAM_MULTI_MAP: mmap_set(&m->acts, m->dumps[v[++i]]); m->has_acts = 1; break;
        case PARAM_CHARMAP: mmapf_set(&m->wbd_charmap, m->dumps[v[++i]]); m->has_wbd_charmap = 1; break;
        case PARAM_ACT_DATA: case PARAM_PUNKT: case PARAM_EOS: case PARAM_EOP: case PARAM_WORD:
        case PARAM_XWORD: case PARAM_SEG: case PARAM_IGNORE: case PARAM_MAX_TAG: ++i; break;
        default: return 0;
        }
    }
    if (fsm && m->has_acts) calc_fn_ini(m);
    if (m->ignore_case) return 0;  /* FAUtf32ToLower tables not restated; no shipped model sets it */
    return fsm != NULL;
}

/* FADictConfKeeper.cpp:57-228 (Init) */
static int init_seg(bfo_model* m, const int* v, int n) {
    int fsm_type = TYPE_MEALY_DFA; m->i2info_mode = MODE_PACK_TRIV;
    for (int i = 0; i < n; ++i) {
        switch (v[i]) {
        case PARAM_NO_TR: break;
        case PARAM_IGNORE_CASE: return 0;
        case PARAM_USE_BYTE_ENCODING: m->use_raw_bytes = 1; break;
        case PARAM_NO_DUMMY_PREFIX: m->no_dummy_prefix = 1; break;
        case PARAM_DIRECTION: ++i; break;
        case PARAM_TOKENIZATION_TYPE: m->tok_algo = v[++i]; break;
        case PARAM_ID_OFFSET: m->id_offset = v[++i]; break;
        case PARAM_FSM_TYPE: fsm_type = v[++i]; break;
        case PARAM_MAP_MODE: m->i2info_mode = v[++i]; break;
        case PARAM_FSM: if (fsm_type != TYPE_MEALY_DFA) return 0;
                        if (!dfa_set(&m->seg_dfa, m->dumps[v[++i]], 1)) return 0;
                        break;
        case PARAM_ARRAY: ++i; break;  /* K2I: identity, never read on this path */
        case PARAM_CHARMAP: mmapf_set(&m->seg_charmap, m->dumps[v[++i]]); m->has_seg_charmap = 1; break;
        case PARAM_MULTI_MAP:
            if (m->i2info_mode == MODE_PACK_FIXED) mmapf_set(&m->i2info_fixed, m->dumps[v[++i]]);
            else if (m->i2info_mode == MODE_PACK_TRIV) mmap_set(&m->i2info_triv, m->dumps[v[++i]]);
            else return 0;
            break;
        default: return 0;
        }
    }
    return 1;
}

static int i2info_get(const bfo_model* m, int key, const int** vals) {
    if (m->i2info_mode == MODE_PACK_FIXED) return mmapf_get_ptr(&m->i2info_fixed, key, vals);
    return mmap_get(&m->i2info_triv, key, vals);
}

void bfo_free_model(bfo_model* m) {
    if (!m) return;
    free(m->fn2ini); free(m->image); free(m);
}

/* blingfiretokdll.cpp:1077-1094 + :918-1048; FAImageDump.cpp:62-118; FALDB.cpp:24-64 */
bfo_model* bfo_load_model(const char* path) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    bfo_model* m = (bfo_model*)calloc(1, sizeof(bfo_model));
    fseek(f, 0, SEEK_END); m->image_size = ftell(f); fseek(f, 0, SEEK_SET);
    m->image = (u8*)malloc((size_t)m->image_size + 16);
    if (m->image_size < 8 || fread(m->image, 1, (size_t)m->image_size, f) != (size_t)m->image_size) { fclose(f); bfo_free_model(m); return NULL; }
    fclose(f);
    int count = rd_i32(m->image);
    if (count <= 0 || count > 255 || 4 + 4L * count > m->image_size) { bfo_free_model(m); return NULL; }
    m->dump_count = count;
    for (int i = 0; i < count; ++i) {
        int off = rd_i32(m->image + 4 + 4 * i);
        if (off < 0 || off >= m->image_size) { bfo_free_model(m); return NULL; }
        m->dumps[i] = m->image + off; m->dump_off[i] = off;
    }
    mmap_set(&m->conf, m->dumps[0]);
    if (!ldb_is_valid(m)) { bfo_free_model(m); return NULL; }
    const int* vals = NULL; int n = mmap_get(&m->conf, FUNC_WBD, &vals);
    if (n != -1) { if (!init_wbd(m, vals, n)) { bfo_free_model(m); return NULL; } m->has_wbd = 1; }
    vals = NULL; n = mmap_get(&m->conf, FUNC_POS_DICT, &vals);
    if (n != -1) { if (!init_seg(m, vals, n)) { bfo_free_model(m); return NULL; } m->has_seg = 1; }
    /* blingfiretokdll.cpp:997-1045 */
    m->min_token_id = 0; m->max_token_id = 1000000000;   /* FALimits::MaxArrSize */
    vals = NULL; n = mmap_get(&m->conf, FUNC_I2W, &vals);
    if (n != -1) {
        for (int i = 0; i < n; ++i) {
            if (vals[i] == PARAM_STRING_ARRAY && i + 1 < n) {
                const u8* d = m->dumps[vals[++i]];
                m->i2w_count = rd_i32(d);                    /* FAStringArray_pack.cpp:22-52 */
                m->i2w_offsets = d + 4; m->i2w_data = d + 4 + 4L * (m->i2w_count + 1);
                m->has_i2w = 1;
            } else if (vals[i] == PARAM_TOKENID_MIN && i + 1 < n) m->min_token_id = vals[++i];
            else if (vals[i] == PARAM_TOKENID_MAX && i + 1 < n) m->max_token_id = vals[++i];
        }
    }
    return m;
}

/* blingfiretokdll.cpp:1669-1679 */
int bfo_set_no_dummy_prefix(bfo_model* m, int flag) { if (!m) return 0; m->no_dummy_prefix = flag ? 1 : 0; return 1; }

/* blingfiretokdll.cpp:1689-1745 over FAStringArray_pack::GetAt (FAStringArray_pack.cpp:55-71) */
int bfo_ids_to_text(const bfo_model* m, const int32_t* ids, int count, char* out, int max_out, int skip_special) {
    if (!m) return 0;
    if (count == 0 || !ids) return 0;
    if (!m->has_i2w) return 0;
    int len = 0;
    for (int i = 0; i < count; ++i) {
        const int id = ids[i];
        if (skip_special && (id < m->min_token_id || id > m->max_token_id)) continue;
        if (id < 0 || id >= m->i2w_count) return 0;
        const unsigned b = rd_u32(m->i2w_offsets + 4L * id), e = rd_u32(m->i2w_offsets + 4L * (id + 1));
        const u8* tok = m->i2w_data + b;
        int tl = (int)(e - b);
        if (len == 0 && tl > 0 && tok[0] == 0x20) { ++tok; --tl; }
        if (tl > 0 && max_out - len >= tl) memcpy(out + len, tok, (size_t)tl);
        len += tl;
    }
    if (max_out > len) out[len] = 0;
    return len + 1;
}

/* ---- UTF-8: FAUtf8Utils.cpp ---- */
static int utf8_size_of_symbol(int s) {  /* :44-57 */
    unsigned u = (unsigned)s;
    if (u <= 0x7F) return 1;
    if (u <= 0x7FF) return 2;
    if (u <= 0xFFFF) return 3;
    if (u <= 0x10FFFF) return 4;
    return 0;
}
static int utf8_size_of_lead(const char* p) {  /* :23-42 */
    int ch = (u8)*p;
    if ((ch & 0x80) == 0) return 1;
    if ((ch & 0xE0) == 0xC0) return 2;
    if ((ch & 0xF0) == 0xE0) return 3;
    if ((ch & 0xF8) == 0xF0) return 4;
    return 0;
}
/* :121-196 (FAUtf8ToInt with end pointer) */
static const char* utf8_to_int(const char* b, const char* e, int* res) {
    if (e <= b) return NULL;
    size_t left = (size_t)(e - b);
    int ch = (u8)*b++, n;
    if ((ch & 0x80) == 0) { *res = ch; return b; }
    if ((ch & 0xE0) == 0xC0) { n = 2; ch &= ~0xE0; }
    else if ((ch & 0xF0) == 0xE0) { n = 3; ch &= ~0xF0; }
    else if ((ch & 0xF8) == 0xF0) { n = 4; ch &= ~0xF8; }
    else return NULL;
    if (left < (unsigned)n) return NULL;
    int r = ch;
    for (int i = 1; i < n; ++i) { r <<= 6; ch = (u8)*b++; if ((ch & 0xC0) != 0x80) return NULL; r |= ch & 0x3f; }
    if (n != utf8_size_of_symbol(r)) return NULL;
    if ((r & 0xFFFFF800) == 0xD800) return NULL;
    *res = r; return b;
}
/* :233-270 / :273-313 (FAStrUtf8ToArray, with optional offsets) */
static int str_utf8_to_array(const char* s, int len, int* arr, int* offs, int max) {
    const char* begin = s; const char* end = s + len; int i = 0;
    if (len >= 3 && (u8)s[0] == 0xEF && (u8)s[1] == 0xBB && (u8)s[2] == 0xBF) s += 3;
    while (s < end && i < max) {
        int off = (int)(s - begin);
        s = utf8_to_int(s, end, arr + i);
        if (!s) return -1;
        if (offs) offs[i] = off;
        i++;
    }
    return i;
}
/* :316-345 / :348-380 (FAStrUtf8AsBytesToArray) */
static int str_utf8_bytes_to_array(const char* s, int len, int* arr, int* offs, int max) {
    const char* begin = s; const char* end = s + len; int i = 0;
    if (len >= 3 && (u8)s[0] == 0xEF && (u8)s[1] == 0xBB && (u8)s[2] == 0xBF) s += 3;
    while (s < end && i < max) { if (offs) offs[i] = (int)(s - begin); arr[i++] = (u8)*s++; }
    return i;
}
/* :471-528 (FAIntToUtf8) */
static char* int_to_utf8(int sym, char* p, int max) {
    unsigned u = (unsigned)sym;
    if (u <= 0x7F && max > 0) { *p++ = (char)u; return p; }
    if (u <= 0x7FF && max > 1) { *p++ = (char)(0xC0 | (u >> 6)); *p++ = (char)(0x80 | (u & 0x3F)); return p; }
    if (u <= 0xFFFF && max > 2) {
        if ((sym & 0xFFFFF800) == 0xD800) return NULL;
        *p++ = (char)(0xE0 | (u >> 12)); *p++ = (char)(0x80 | ((u >> 6) & 0x3F)); *p++ = (char)(0x80 | (u & 0x3F)); return p;
    }
    if (u <= 0x10FFFF && max > 3) {
        *p++ = (char)(0xF0 | (u >> 18)); *p++ = (char)(0x80 | ((u >> 12) & 0x3F));
        *p++ = (char)(0x80 | ((u >> 6) & 0x3F)); *p++ = (char)(0x80 | (u & 0x3F)); return p;
    }
    return NULL;
}

/* FAUtils_cl.h:311-369 / :372-440 (FANormalize, with optional offsets) */
static int normalize(const int* in, int n, int* out, int* offs, int max_out, const mmap_fixed_t* map) {
    int norm[10]; int o = 0;
    for (int i = 0; i < n; ++i) {
        int ci = in[i]; int cnt = mmapf_get_copy(map, ci, norm, 10);
        if (cnt == -1) { if (o < max_out) { out[o] = ci; if (offs) offs[o] = i; } o++; }
        else if (cnt == 1) { if (o < max_out) { out[o] = norm[0]; if (offs) offs[o] = i; } o++; }
        else if (cnt > 1 && cnt <= 10) {
            int cc = max_out - o; if (cnt < cc) cc = cnt;
            for (int j = 0; j < cc; ++j) { out[o + j] = norm[j]; if (offs) offs[o + j] = i; }
            o += cnt;
        }
    }
    return o;
}

/* FALexTools_t.h:205-400 (Process_int) */
static int lex_process_int(const bfo_model* m, int initial, int offset, const int* in, int n,
                           int* out, int max_out, int depth, int once) {
    const dfa_t* d = &m->wbd_dfa; int out_size = 0;
    if (m->max_depth < depth) return 0;
    for (int from = -1; from < n; ++from) {
        int state = initial, fstate = -1, fpos = -1, j = from, dst;
        int bound = from + m->max_token_length; if (n < bound) bound = n;
        if (j == -1) {
            state = dfa_get_dest(d, initial, IW_L_ANCHOR);
            if (state == -1) { state = dfa_get_dest(d, initial, IW_ANY); if (state == -1) continue; }
            j++;
        }
        for (; j < bound; ++j) {
            int iw = in[j]; if (iw < IW_EPSILON) iw = IW_EPSILON;
            dst = dfa_get_dest(d, state, iw);
            if (dst == -1) { dst = dfa_get_dest(d, state, IW_ANY); if (dst == -1) break; }
            if (dfa_is_final(d, dst)) { fstate = dst; fpos = j; }
            state = dst;
        }
        if (j == n) {
            dst = dfa_get_dest(d, state, IW_R_ANCHOR);
            if (dst == -1) dst = dfa_get_dest(d, state, IW_ANY);
            if (dst != -1 && dfa_is_final(d, dst)) { fstate = dst; fpos = j; }
        }
        if (fpos == -1) continue;
        const int ow = dfa_get_ow(d, fstate);
        const int* act = NULL; const int act_size = mmap_get(&m->acts, ow, &act);
        if (act_size < 3 || !act) return out_size;  /* reference asserts; malformed model */
        const int left = act[0], right = act[1], tag = act[2];
        int from2 = from + left; if (from2 < 0) from2 = 0; else if (n <= from2) from2 = n - 1;
        int to2 = fpos - right; if (to2 < 0) to2 = 0; else if (n <= to2) to2 = n - 1;
        int fn_idx = 3;
        if (tag != 0) {
            if (out_size + 3 <= max_out) { out[out_size++] = tag; out[out_size++] = from2 + offset; out[out_size++] = to2 + offset; }
            else return out_size;
            fn_idx = 4;
        }
        const int fn_once = 1 < (act_size - fn_idx);
        int fn_from = from2;
        for (; fn_idx < act_size; ++fn_idx) {
            const int fn = act[fn_idx];
            if (fn < 0 || fn >= m->fn2ini_size) break;
            const int r = lex_process_int(m, m->fn2ini[fn], fn_from + offset, in + fn_from, to2 - fn_from + 1,
                                          out + out_size, max_out - out_size, depth + 1, fn == 0 ? 0 : fn_once);
            if (r > 0) { out_size += r; fn_from = out[out_size - 1] + 1 - offset; if (fn_from > to2) break; }
        }
        if (once) return out_size;
        if (fpos - right > from) from = fpos - right;
    }
    return out_size;
}

int bfo_lex_process(const bfo_model* m, const int* in, int n, int* out, int max_out) {  /* FALexTools_t.h:403-421 */
    if (!m || !m->has_wbd || !m->has_acts) return -1;
    return lex_process_int(m, m->wbd_dfa.initial, 0, in, n, out, max_out, 1, 0);
}

/* blingfiretokdll.cpp:1108-1314 (TextToIdsWithOffsets_wp) */
static int text_to_ids_wp(const bfo_model* m, const char* s, int n, int32_t* ids, int* starts, int* ends, int max_ids, int unk) {
    if (n <= 0 || n > MAX_ARR_SIZE || !s || !m) return 0;
    const int need_offs = starts && ends;
    int* buf = (int*)malloc(sizeof(int) * (size_t)n * 4);
    int* offs = buf + n; int* nbuf = buf + 2 * (size_t)n; int* noffs = buf + 3 * (size_t)n;
    int ret = 0, size = str_utf8_to_array(s, n, buf, need_offs ? offs : NULL, n);
    const int* in = buf;
    if (size <= 0 || size > n) goto done;
    if (m->has_wbd_charmap) {
        size = normalize(buf, size, nbuf, need_offs ? noffs : NULL, n, &m->wbd_charmap);
        if (size <= 0 || size > n) goto done;
        in = nbuf;
    }
    {
        const int max_res = size * 6;
        int* res = (int*)malloc(sizeof(int) * (size_t)max_res);
        const int rn = bfo_lex_process(m, in, size, res, max_res);
        if (rn > max_res || rn % 3 != 0 || rn < 0) { free(res); goto done; }
        int out = 0;
        for (int i = 0; i < rn; i += 3) {
            const int tag = res[i];
            if (tag == WBD_IGNORE_TAG) continue;
            if (tag == WBD_WORD_TAG) {
                const int tfrom = res[i + 1], tto = res[i + 2];
                int j = i + 3, nsub = 0, covered = 0;
                if (j < rn) {
                    int exp = tfrom, stag = res[j], sfrom = res[j + 1], sto = res[j + 2];
                    while (j <= rn && stag > WBD_IGNORE_TAG && exp == sfrom) {
                        exp = sto + 1; nsub++; j += 3;
                        if (j < rn) { stag = res[j]; sfrom = res[j + 1]; sto = res[j + 2]; }
                    }
                    if (nsub > 0 && exp - 1 == tto) {
                        for (int k = 0; k < nsub && out < max_ids; ++k) {
                            const int ti = (k + 1) * 3 + i;
                            ids[out] = res[ti];
                            if (need_offs) {
                                const int sf = res[ti + 1], st = res[ti + 2];
                                starts[out] = offs[m->has_wbd_charmap ? noffs[sf] : sf];
                                const int to_off = offs[m->has_wbd_charmap ? noffs[st] : st];
                                const int cs = utf8_size_of_lead(s + to_off);
                                ends[out] = to_off + (cs > 0 ? cs - 1 : 0);
                            }
                            out++;
                        }
                        covered = 1;
                    }
                }
                if (!covered && out < max_ids) {
                    ids[out] = unk;
                    if (need_offs) {
                        starts[out] = offs[m->has_wbd_charmap ? noffs[tfrom] : tfrom];
                        const int to_off = offs[m->has_wbd_charmap ? noffs[tto] : tto];
                        const int cs = utf8_size_of_lead(s + to_off);
                        ends[out] = to_off + (cs > 0 ? cs - 1 : 0);
                    }
                    out++;
                }
                i = j - 3;
            }
            if (out >= max_ids) break;
        }
        ret = out;
        free(res);
    }
done:
    free(buf);
    return ret;
}

static int is_white(int c) {  /* blingfiretokdll.h:17-21 (__FAIsWhiteSpace__) */
    return c <= 0x20 || c == 0xa0 || (c >= 0x2000 && c <= 0x200f) || c == 0x202f || c == 0x205f ||
           c == 0x2060 || c == 0x2420 || c == 0x2424 || c == 0x3000 || c == 0xfeff;
}

/* ---- segmentation algorithms over the Mealy MPH automaton ---- */
typedef struct { int begin, id; double score; } uni_arc;

/* FATokenSegmentationTools_1best_t.h:174-279 (Unigram-LM best path) */
static int seg_unigram(const bfo_model* m, const int* in, int n, int* out, int max_out, int unk) {
    if (n <= 0) return 0;
    const dfa_t* d = &m->seg_dfa;
    uni_arc* arcs = (uni_arc*)malloc(sizeof(uni_arc) * (size_t)n);
    for (int i = 0; i < n; ++i) { arcs[i].begin = -1; arcs[i].id = -1; arcs[i].score = -FLT_MAX; }
    const float unk_score = -100000.0f;
    for (int start = 0; start < n; ++start) {
        int state = d->initial, sum = 0, ow = 0, unknown = 1;
        for (int i = start; i < n; ++i) {
            state = mealy_get_dest_ow(d, state, in[i], &ow);
            if (state == -1) break;
            sum += ow;
            if (dfa_is_final(d, state)) {  /* AddArc :118-142 */
                const int* v = NULL; int c = i2info_get(m, sum, &v);
                if (c != 2 || !v) { free(arcs); return 0; }
                float score; memcpy(&score, &v[1], 4);
                const double prev = start > 0 ? arcs[start - 1].score : 0;
                uni_arc* a = arcs + i;
                if (a->score < score + prev) { a->begin = start; a->id = v[0]; a->score = score + prev; }
                unknown = 0;
            }
        }
        if (unknown) {  /* AddUnknownArc :145-171 */
            uni_arc* a = arcs + start; uni_arc* p = a - 1;
            const double prev = start > 0 ? p->score : 0;
            if (a->score < unk_score + prev) {
                a->begin = start; a->id = -1; a->score = unk_score + prev;
                if (start > 0 && p->id == -1) a->begin = p->begin;
            }
        }
    }
    int actual = 0, end = n - 1;
    while (end >= 0) {
        const uni_arc* a = arcs + end;
        if (actual + 3 <= max_out) { out[actual] = end; out[actual + 1] = a->begin; out[actual + 2] = a->id != -1 ? a->id : unk; }
        actual += 3;
        end = a->begin - 1;
    }
    if (max_out >= actual) for (int i = 0; i < actual / 2; ++i) { int t = out[i]; out[i] = out[actual - i - 1]; out[actual - i - 1] = t; }
    free(arcs);
    return actual;
}

typedef struct { int start, end, id; float rank; } bpe_arc;
static int cmp_bpe(const void* a, const void* b) {  /* ..._bpe_t.h:238-255 */
    const bpe_arc* x = (const bpe_arc*)a; const bpe_arc* y = (const bpe_arc*)b;
    if (x->id < y->id) return -1;
    if (x->id == y->id) { if (x->start < y->start) return -1; if (x->start == y->start) return 0; return 1; }
    return 1;
}
static int cmp_bpe_merges(const void* a, const void* b) {  /* ..._bpe_with_merges_t.h:242-262 */
    const bpe_arc* x = (const bpe_arc*)a; const bpe_arc* y = (const bpe_arc*)b;
    if (x->rank > y->rank) return -1;
    if (x->rank == y->rank) return cmp_bpe(a, b);
    return 1;
}
/* FATokenSegmentationTools_1best_bpe_t.h:125-316 and ..._bpe_with_merges_t.h (same flow, other sort key) */
static int seg_bpe(const bfo_model* m, const int* in, int n, int* out, int max_out, int unk) {
    if (n <= 0) return 0;
    const dfa_t* d = &m->seg_dfa;
    const int with_merges = m->tok_algo == TOKENIZE_BPE_OPT_WITH_MERGES;
    const int fast = with_merges || m->tok_algo == TOKENIZE_BPE_OPT;
    size_t cap = (size_t)n + 16, cnt = 0;
    bpe_arc* arcs = (bpe_arc*)malloc(sizeof(bpe_arc) * cap);
    for (int start = 0; start < n; ++start) {
        int state = d->initial, sum = 0, ow = 0, unknown = 1;
        const int tok_start = in[start] == SP_DELIM;
        const size_t cnt0 = cnt; int ff = start;
        for (int i = start; i < n; ++i) {
            state = mealy_get_dest_ow(d, state, in[i], &ow);
            if (state == -1) break;
            sum += ow;
            if (dfa_is_final(d, state)) {
                const int* v = NULL; int c = i2info_get(m, sum, &v);
                if (c < 1 || !v) { free(arcs); return 0; }
                float rank = 0.0f; if (with_merges) memcpy(&rank, &v[1], 4);
                const int opt = fast && tok_start && ((i < n - 1) ? in[i + 1] == SP_DELIM : 1) && cnt0 < cnt;
                if (cnt + 1 > cap) { cap *= 2; arcs = (bpe_arc*)realloc(arcs, sizeof(bpe_arc) * cap); }
                if (!opt) { arcs[cnt].start = start; arcs[cnt].end = i; arcs[cnt].id = v[0]; arcs[cnt].rank = rank; cnt++; }
                else { arcs[cnt0].start = start; arcs[cnt0].end = i; arcs[cnt0].id = v[0]; arcs[cnt0].rank = rank; cnt = cnt0 + 1; ff = i; }
                unknown = 0;
            }
        }
        if (unknown) {
            if (cnt > 0 && arcs[cnt - 1].id == unk) arcs[cnt - 1].end = start;
            else {
                if (cnt + 1 > cap) { cap *= 2; arcs = (bpe_arc*)realloc(arcs, sizeof(bpe_arc) * cap); }
                arcs[cnt].start = start; arcs[cnt].end = start; arcs[cnt].id = unk; arcs[cnt].rank = 0.0f; cnt++;
            }
        }
        if (fast) start = ff;
    }
    qsort(arcs, cnt, sizeof(bpe_arc), with_merges ? cmp_bpe_merges : cmp_bpe);
    int* tos = (int*)calloc((size_t)n * 2, sizeof(int)); int* ids = tos + n;
    u8* inter = (u8*)calloc((size_t)n + 1, 1);
    for (int i = 0; i < n; ++i) ids[i] = unk;
    for (size_t i = 0; i < cnt; ++i) {
        const int st = arcs[i].start, en = arcs[i].end;
        if (inter[st] == 0 && (en + 1 == n || inter[en + 1] == 0)) {
            tos[st] = en; ids[st] = arcs[i].id;
            if (en - st > 0) memset(inter + st + 1, 1, (size_t)(en - st));
        }
    }
    int actual = 0;
    for (int start = 0; start < n; start++) {
        const int end = tos[start];
        if (actual + 3 <= max_out) { out[actual] = ids[start]; out[actual + 1] = start; out[actual + 2] = end; }
        actual += 3;
        start = end;
    }
    free(arcs); free(tos); free(inter);
    return actual;
}

/* blingfiretokdll.cpp:1349-1535 (TextToIdsWithOffsets_sp) */
static int text_to_ids_sp(const bfo_model* m, const char* s, int n, int32_t* ids, int* starts, int* ends, int max_ids, int unk) {
    if (n <= 0 || n > MAX_ARR_SIZE || !s || !m) return 0;
    const int need_offs = starts && ends;
    const size_t nn = (size_t)n + 1;
    int* buf = (int*)malloc(sizeof(int) * nn * 6);
    int* offs = buf + nn; int* nbuf = buf + 2 * nn; int* noffs = buf + 4 * nn;
    int ret = 0;
    buf[0] = SP_DELIM; offs[0] = -1;
    const int data_off = m->no_dummy_prefix ? 0 : 1;
    int size = m->use_raw_bytes ? str_utf8_bytes_to_array(s, n, buf + data_off, need_offs ? offs + data_off : NULL, n)
                                : str_utf8_to_array(s, n, buf + data_off, need_offs ? offs + data_off : NULL, n);
    if (size <= 0 || size > n) { free(buf); return 0; }
    size += data_off;
    int* cur = buf; int use_norm = 0;
    if (m->has_seg_charmap) {
        const int max_norm = (n + 1) * 2;
        const int ns = normalize(buf, size, nbuf, need_offs ? noffs : NULL, max_norm, &m->seg_charmap);
        if (ns <= 0 || ns > max_norm) { free(buf); return 0; }
        size = ns; cur = nbuf; use_norm = 1;
    }
    int* adj = need_offs ? (use_norm ? noffs : offs) : NULL;
    int i = 0, j = 0;
    while (i < size) {
        const int c = cur[i];
        if (!is_white(c)) { cur[j] = c; if (adj) adj[j] = adj[i]; j++; }
        else if (j == 0 || cur[j - 1] != SP_DELIM) { cur[j] = SP_DELIM; if (adj) adj[j] = adj[i]; j++; }
        i++;
    }
    if (j > 1 && cur[j - 1] == SP_DELIM) j--;
    size = j;
    {
        const int max_res = size * 3;
        int* res = (int*)malloc(sizeof(int) * (size_t)(max_res > 0 ? max_res : 1));
        int rn;
        if (m->tok_algo == TOKENIZE_BPE || m->tok_algo == TOKENIZE_BPE_OPT || m->tok_algo == TOKENIZE_BPE_OPT_WITH_MERGES)
            rn = seg_bpe(m, cur, size, res, max_res, unk);
        else
            rn = seg_unigram(m, cur, size, res, max_res, unk);
        if (rn <= max_res && rn % 3 == 0) {
            int out = 0;
            for (int k = 0; k < rn && out < max_ids; k += 3) {
                ids[out] = res[k] + m->id_offset;
                if (need_offs) {
                    const int tf = res[k + 1], tt = res[k + 2];
                    starts[out] = offs[use_norm ? noffs[tf] : tf];
                    const int to_off = offs[use_norm ? noffs[tt] : tt];
                    /* to_off == -1 for a token that is only the dummy prefix: the reference then
                     * reads the byte BEFORE the input (blingfiretokdll.cpp:1527, out of bounds, value
                     * unspecified).  The oracle pins that case to size 0. */
                    const int cs = to_off < 0 ? 0 : utf8_size_of_lead(s + to_off);
                    ends[out] = to_off + (cs > 0 ? cs - 1 : 0);
                }
                out++;
            }
            ret = out;
        }
        free(res);
    }
    free(buf);
    return ret;
}

int bfo_text_to_ids_with_offsets(const bfo_model* m, const char* s, int n, int32_t* ids, int* starts, int* ends, int max_ids, int unk) {
    if (!m) return 0;  /* blingfiretokdll.cpp:1563-1609 */
    if (!m->has_seg) { if (!m->has_wbd) return 0; return text_to_ids_wp(m, s, n, ids, starts, ends, max_ids, unk); }
    return text_to_ids_sp(m, s, n, ids, starts, ends, max_ids, unk);
}
int bfo_text_to_ids(const bfo_model* m, const char* s, int n, int32_t* ids, int max_ids, int unk) {
    return bfo_text_to_ids_with_offsets(m, s, n, ids, NULL, NULL, max_ids, unk);  /* :1619-1646 */
}

/* blingfiretokdll.cpp:415-566 (TextToWordsWithOffsetsWithModel); starts/ends may be NULL */
int bfo_text_to_words_with_offsets(const bfo_model* m, const char* s, int n, char* out, int* starts, int* ends, int max_out) {
    if (!m || !m->has_wbd) return -1;
    if (n == 0) return 0;
    if (n < 0 || n > MAX_ARR_SIZE || !s) return -1;
    int* buf = (int*)malloc(sizeof(int) * (size_t)n * 5);
    int* offs = buf + n; int* res = buf + 2 * (size_t)n;
    int ret = -1;
    if (starts && max_out > 0) memset(starts, 0, sizeof(int) * (size_t)max_out);   /* :467-472 */
    if (ends && max_out > 0) memset(ends, 0, sizeof(int) * (size_t)max_out);
    const int size = str_utf8_to_array(s, n, buf, offs, n);
    if (size <= 0 || size > n) { free(buf); return -1; }
    for (int i = 0; i < size; ++i) if (buf[i] == 0) buf[i] = 0x20;
    const int rn = bfo_lex_process(m, buf, size, res, size * 3);
    if (rn > size * 3 || rn % 3 != 0 || rn < 0) { free(buf); return -1; }
    char* acc = (char*)malloc((size_t)n * 2 + 16); size_t len = 0; int added = 0, words = 0;
    for (int i = 0; i < rn; i += 3) {
        if (res[i] == WBD_IGNORE_TAG) continue;
        const int from = res[i + 1], to = res[i + 2];
        if (starts && words < max_out) starts[words] = offs[from];                 /* :523-525 */
        if (ends && words < max_out) { const int cs = utf8_size_of_lead(s + offs[to]); ends[words] = offs[to] + (cs > 0 ? cs - 1 : 0); }
        words++;
        if (added) acc[len++] = ' ';
        char* p = acc + len; char* q = p; int budget = n;
        for (int k = from; k <= to; ++k) {
            char* nx = int_to_utf8(buf[k], q, budget - (int)(q - p));
            if (!nx) { free(acc); free(buf); return -1; }
            q = nx;
        }
        for (char* c = p; c < q; ++c) if (*c == ' ') *c = '_';
        /* the reference appends the token as a C string: it stops at an embedded NUL (:550) */
        size_t tl = 0; while (p + tl < q && p[tl] != 0) tl++;
        len += tl; added = 1;
    }
    acc[len++] = 0;
    ret = (int)len;
    if (ret <= max_out && out) memcpy(out, acc, len);
    free(acc); free(buf);
    return ret;
}
int bfo_text_to_words(const bfo_model* m, const char* s, int n, char* out, int max_out) {
    return bfo_text_to_words_with_offsets(m, s, n, out, NULL, NULL, max_out);
}

/* blingfiretokdll.cpp:163-355 (TextToSentencesWithOffsetsWithModel); starts/ends may be NULL.
 * One sentence per triple: From = previous To + 1 (tags and Froms of the triples are ignored), leading
 * white space skipped, '\n' inside a sentence -> ' ', sentences joined by '\n', the rest of the
 * paragraph after the last boundary is the last sentence. */
static size_t sentence_append(const char* s, const int* buf, const int* offs, int from, int to, int n, char* acc, size_t len,
                              int* added, int* count, int* starts, int* ends, int max_out, int* err) {
    int delta = 0;
    while (delta < to - from + 1 && is_white(buf[from + delta])) delta++;           /* FAGetFirstNonWhiteSpace :138-150 */
    if (delta >= to - from + 1) return len;
    if (starts && *count < max_out) starts[*count] = offs[from + delta];
    if (ends && *count < max_out) { const int cs = utf8_size_of_lead(s + offs[to]); ends[*count] = offs[to] + (cs > 0 ? cs - 1 : 0); }
    (*count)++;
    if (*added) acc[len++] = '\n';
    char* p = acc + len; char* q = p;
    for (int k = from + delta; k <= to; ++k) {
        char* nx = int_to_utf8(buf[k], q, n - (int)(q - p));
        if (!nx) { *err = 1; return len; }
        q = nx;
    }
    for (char* c = p; c < q; ++c) if (*c == '\n') *c = ' ';
    size_t tl = 0; while (p + tl < q && p[tl] != 0) tl++;                             /* appended as a C string */
    *added = 1;
    return len + tl;
}
int bfo_text_to_sentences_with_offsets(const bfo_model* m, const char* s, int n, char* out, int* starts, int* ends, int max_out) {
    if (!m || !m->has_wbd) return -1;
    if (n == 0) return 0;
    if (n < 0 || n > MAX_ARR_SIZE || !s) return -1;
    int* buf = (int*)malloc(sizeof(int) * (size_t)n * 5);
    int* offs = buf + n; int* res = buf + 2 * (size_t)n;
    if (starts && max_out > 0) memset(starts, 0, sizeof(int) * (size_t)max_out);
    if (ends && max_out > 0) memset(ends, 0, sizeof(int) * (size_t)max_out);
    const int size = str_utf8_to_array(s, n, buf, offs, n);
    if (size <= 0 || size > n) { free(buf); return -1; }
    for (int i = 0; i < size; ++i) if (buf[i] == 0) buf[i] = 0x20;
    const int rn = bfo_lex_process(m, buf, size, res, size * 3);
    if (rn > size * 3 || rn % 3 != 0 || rn < 0) { free(buf); return -1; }
    char* acc = (char*)malloc((size_t)n * 2 + 16); size_t len = 0; int added = 0, count = 0, err = 0, prev_end = -1;
    for (int i = 0; i < rn && !err; i += 3) {
        const int from = prev_end + 1, to = res[i + 2];
        prev_end = to;
        len = sentence_append(s, buf, offs, from, to, n, acc, len, &added, &count, starts, ends, max_out, &err);
    }
    if (!err && prev_end + 1 < size)                                                  /* :303-338 */
        len = sentence_append(s, buf, offs, prev_end + 1, size - 1, n, acc, len, &added, &count, starts, ends, max_out, &err);
    int ret = -1;
    if (!err) {
        acc[len++] = 0;
        ret = (int)len;
        if (ret <= max_out && out) memcpy(out, acc, len);
    }
    free(acc); free(buf);
    return ret;
}

/* ---- introspection ---- */
int bfo_dfa_initial(const bfo_model* m) { return m->has_seg ? m->seg_dfa.initial : m->wbd_dfa.initial; }
int bfo_dfa_get_dest(const bfo_model* m, int s, int iw) { return dfa_get_dest(&m->wbd_dfa, s, iw); }
int bfo_dfa_is_final(const bfo_model* m, int s) { return dfa_is_final(m->has_wbd ? &m->wbd_dfa : &m->seg_dfa, s); }
int bfo_dfa_get_ow(const bfo_model* m, int s) { return dfa_get_ow(&m->wbd_dfa, s); }
int bfo_iwmap_new_iw(const bfo_model* m, int iw) { return m->wbd_dfa.remap ? iwmap_get(&m->wbd_dfa.iwmap, iw) : iw; }
int bfo_act_get(const bfo_model* m, int key, const int** vals) { return mmap_get(&m->acts, key, vals); }
int bfo_charmap_get(const bfo_model* m, int cp, int* out, int max_out) {
    const mmap_fixed_t* c = m->has_seg ? (m->has_seg_charmap ? &m->seg_charmap : NULL) : (m->has_wbd_charmap ? &m->wbd_charmap : NULL);
    if (!c) return -1;
    return mmapf_get_copy(c, cp, out, max_out);
}
int bfo_fn_ini(const bfo_model* m, int fn) { return (fn >= 0 && fn < m->fn2ini_size) ? m->fn2ini[fn] : -1; }
int bfo_has_seg(const bfo_model* m) { return m->has_seg; }

/* bulk variants so the table cross-checks in tests/ do not pay one ctypes call per lookup */
void bfo_dfa_get_dest_many(const bfo_model* m, const int* states, const int* iws, long n, int* out) {
    for (long i = 0; i < n; ++i) out[i] = dfa_get_dest(&m->wbd_dfa, states[i], iws[i]);
}
void bfo_dfa_row(const bfo_model* m, int state, const int* iws, int n, int* out) {
    for (int i = 0; i < n; ++i) out[i] = dfa_get_dest(&m->wbd_dfa, state, iws[i]);
}
void bfo_iwmap_many(const bfo_model* m, int from, int to, int* out) {
    for (int iw = from; iw < to; ++iw) out[iw - from] = m->wbd_dfa.remap ? iwmap_get(&m->wbd_dfa.iwmap, iw) : iw;
}
void bfo_state_info_many(const bfo_model* m, const int* states, long n, int* finals, int* ows) {
    for (long i = 0; i < n; ++i) { finals[i] = dfa_is_final(&m->wbd_dfa, states[i]); ows[i] = finals[i] ? dfa_get_ow(&m->wbd_dfa, states[i]) : -1; }
}
/* class of a code point as the lexer sees it: charmap (1->1 rows), clamp, class map */
void bfo_lexer_class_many(const bfo_model* m, int from, int to, int* out) {
    int norm[10];
    for (int cp = from; cp < to; ++cp) {
        int x = cp;
        if (m->has_wbd_charmap) { int c = mmapf_get_copy(&m->wbd_charmap, cp, norm, 10); if (c == 1) x = norm[0]; else if (c != -1) { out[cp - from] = -2; continue; } }
        if (x < IW_EPSILON) x = IW_EPSILON;
        out[cp - from] = m->wbd_dfa.remap ? iwmap_get(&m->wbd_dfa.iwmap, x) : x;
    }
}

/* ---- threaded batch driver (CPU baseline timing) ---- */
typedef struct { const bfo_model* m; const char* utf8; const int64_t* offs; int64_t ndocs; int32_t* ids; int32_t* counts;
                 int max_ids, unk, tid, nthreads; int64_t total; } batch_job;
static void* batch_worker(void* arg) {
    batch_job* j = (batch_job*)arg; int64_t tot = 0;
    for (int64_t d = j->tid; d < j->ndocs; d += j->nthreads) {
        const int64_t b = j->offs[d], e = j->offs[d + 1];
        const int c = bfo_text_to_ids(j->m, j->utf8 + b, (int)(e - b), j->ids + d * j->max_ids, j->max_ids, j->unk);
        j->counts[d] = c; tot += c;
    }
    j->total = tot; return NULL;
}
int64_t bfo_text_to_ids_batch(const bfo_model* m, const char* utf8, const int64_t* offsets, int64_t ndocs,
                              int32_t* ids, int32_t* counts, int max_ids, int unk, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    batch_job jobs[256]; pthread_t th[256]; int64_t total = 0;
    for (int t = 0; t < threads; ++t) {
        batch_job j = { m, utf8, offsets, ndocs, ids, counts, max_ids, unk, t, threads, 0 }; jobs[t] = j;
        pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) { pthread_join(th[t], NULL); total += jobs[t].total; }
    return total;
}

/* ---- per-document digests for full-size parity (bench.py, tests/): FNV-1a-64 over the uint32 id
 * stream of a document (SURVEY 8c recipe), computed without materialising an [ndocs][max_ids] matrix ---- */
static uint64_t fnv_ids(const int32_t* ids, int n) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (int i = 0; i < n; ++i) { h ^= (uint32_t)ids[i]; h *= 0x100000001b3ull; }
    return h;
}
typedef struct { const bfo_model* m; const char* utf8; const int64_t* offs; int64_t ndocs; uint64_t* dig; int32_t* counts;
                 int max_ids, unk, tid, nthreads; int64_t total; } dig_job_t;
static void* dig_worker(void* arg) {
    dig_job_t* j = (dig_job_t*)arg;
    int32_t* ids = (int32_t*)malloc(sizeof(int32_t) * (size_t)(j->max_ids > 0 ? j->max_ids : 1));
    int64_t tot = 0;
    for (int64_t d = j->tid; d < j->ndocs; d += j->nthreads) {
        int c = bfo_text_to_ids(j->m, j->utf8 + j->offs[d], (int)(j->offs[d + 1] - j->offs[d]), ids, j->max_ids, j->unk);
        if (c > j->max_ids) c = j->max_ids;
        j->counts[d] = c; j->dig[d] = fnv_ids(ids, c); tot += c;
    }
    j->total = tot; free(ids);
    return NULL;
}
/* digests[d], counts[d] of the oracle's TextToIds on every document; returns the total id count */
int64_t bfo_text_to_ids_digests(const bfo_model* m, const char* utf8, const int64_t* offsets, int64_t ndocs,
                                uint64_t* digests, int32_t* counts, int max_ids, int unk_id, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 512) threads = 512;
    dig_job_t* jobs = (dig_job_t*)calloc((size_t)threads, sizeof(dig_job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; ++t) {
        dig_job_t j = { m, utf8, offsets, ndocs, digests, counts, max_ids, unk_id, t, threads, 0 };
        jobs[t] = j; pthread_create(&th[t], NULL, dig_worker, &jobs[t]);
    }
    int64_t tot = 0;
    for (int t = 0; t < threads; ++t) { pthread_join(th[t], NULL); tot += jobs[t].total; }
    free(jobs); free(th);
    return tot;
}
/* the same digests of ids given in CSR form (what the product's batch call returned); elem_size 4 or 2 */
typedef struct { const void* ids; int elem; const int64_t* off; int64_t ndocs; uint64_t* dig; int tid, nthreads; } csr_job_t;
static void* csr_worker(void* arg) {
    csr_job_t* j = (csr_job_t*)arg;
    for (int64_t d = j->tid; d < j->ndocs; d += j->nthreads) {
        uint64_t h = 0xcbf29ce484222325ull;
        for (int64_t k = j->off[d]; k < j->off[d + 1]; ++k) {
            const uint32_t v = j->elem == 2 ? (uint32_t)((const uint16_t*)j->ids)[k] : (uint32_t)((const int32_t*)j->ids)[k];
            h ^= v; h *= 0x100000001b3ull;
        }
        j->dig[d] = h;
    }
    return NULL;
}
void bfo_csr_digests(const void* ids, int elem_size, const int64_t* id_offsets, int64_t ndocs, uint64_t* digests, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 512) threads = 512;
    csr_job_t* jobs = (csr_job_t*)calloc((size_t)threads, sizeof(csr_job_t));
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; ++t) {
        csr_job_t j = { ids, elem_size, id_offsets, ndocs, digests, t, threads };
        jobs[t] = j; pthread_create(&th[t], NULL, csr_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], NULL);
    free(jobs); free(th);
}
/* one number for a whole batch: FNV-1a-64 over the per-document (count, digest) pairs in document order */
uint64_t bfo_fold_digests(const uint64_t* digests, const int64_t* id_offsets, const int32_t* counts, int64_t ndocs) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (int64_t d = 0; d < ndocs; ++d) {
        const uint64_t c = id_offsets ? (uint64_t)(id_offsets[d + 1] - id_offsets[d]) : (uint64_t)counts[d];
        h ^= c; h *= 0x100000001b3ull;
        h ^= digests[d]; h *= 0x100000001b3ull;
    }
    return h;
}
