/*
 * bf_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's TextToIds / TextToWords hot path
 * (microsoft/BlingFire).  It reads the packed .bin LDB image directly, the way
 * the reference readers do, and is deliberately independent of the product's
 * flattened tables so that the two can be cross-checked.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this.  The product library (blingfire_b200/) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks it against
 *   (a) the known-answer vectors the reference documents (README.md:111,131-135,
 *       blingfiretokdll.cpp:1104-1106, README.md:267-268) and
 *   (b) the reference itself compiled from its own sources into oracle/_ref/
 *       (oracle/Makefile), doc by doc on the reference's own corpora, with the
 *       golden digests committed under tests/golden/.
 */
#ifndef BF_ORACLE_H
#define BF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bfo_model bfo_model;

/* blingfiretokdll.cpp:1077-1094 (LoadModel) + :918-1048 (SetModelData). NULL on failure. */
bfo_model* bfo_load_model(const char* path);
void bfo_free_model(bfo_model* m);

/* blingfiretokdll.cpp:1619-1646 (TextToIds): dispatches on [pos-dict] presence. */
int bfo_text_to_ids(const bfo_model* m, const char* utf8, int nbytes,
                    int32_t* ids, int max_ids, int unk_id);

/* blingfiretokdll.cpp:1669-1679 (SetNoDummyPrefix) and :1689-1745 (IdsToText) */
int bfo_set_no_dummy_prefix(bfo_model* m, int flag);
int bfo_ids_to_text(const bfo_model* m, const int32_t* ids, int count, char* out, int max_out, int skip_special);

/* blingfiretokdll.cpp:1108-1314 / :1349-1535 with offsets (NULL to skip). */
int bfo_text_to_ids_with_offsets(const bfo_model* m, const char* utf8, int nbytes,
                                 int32_t* ids, int* starts, int* ends,
                                 int max_ids, int unk_id);

/* blingfiretokdll.cpp:415-566 (TextToWordsWithOffsetsWithModel, offsets omitted). */
int bfo_text_to_words(const bfo_model* m, const char* utf8, int nbytes,
                      char* out, int max_out);
/* blingfiretokdll.cpp:415-566 with the offset arrays (either may be NULL; both are zero-filled first, :467-472) */
int bfo_text_to_words_with_offsets(const bfo_model* m, const char* utf8, int nbytes, char* out, int* starts, int* ends, int max_out);
/* blingfiretokdll.cpp:163-355 (TextToSentencesWithOffsetsWithModel) over a sentence-breaking [wbd] model (sbd.bin) */
int bfo_text_to_sentences_with_offsets(const bfo_model* m, const char* utf8, int nbytes, char* out, int* starts, int* ends, int max_out);

/* FALexTools_t.h:403-421 (Process): raw (Tag,From,To) triples over UTF-32 input. */
int bfo_lex_process(const bfo_model* m, const int* in, int n, int* out, int max_out);

/* introspection used by the table cross-checks in tests/ */
int bfo_dfa_initial(const bfo_model* m);
int bfo_dfa_get_dest(const bfo_model* m, int state, int iw);   /* FARSDfa_pack_triv.cpp:141-399 */
int bfo_dfa_is_final(const bfo_model* m, int state);           /* :128-138 */
int bfo_dfa_get_ow(const bfo_model* m, int state);             /* FAState2Ow_pack_triv.cpp:34-130 */
int bfo_iwmap_new_iw(const bfo_model* m, int iw);              /* FAIwMap_pack.h:55-109 */
int bfo_act_get(const bfo_model* m, int key, const int** vals);/* FAMultiMap_pack.cpp:106-126 */
int bfo_charmap_get(const bfo_model* m, int cp, int* out, int max_out); /* FAMultiMap_pack_fixed.cpp:67-137 */
int bfo_fn_ini(const bfo_model* m, int fn);                    /* FAWbdConfKeeper.cpp:246-314 */
int bfo_has_seg(const bfo_model* m);

/* Batch driver for the CPU baseline: loops bfo_text_to_ids over a CSR batch on
 * `threads` host threads (docs strided).  Returns total ids produced. */
int64_t bfo_text_to_ids_batch(const bfo_model* m, const char* utf8, const int64_t* offsets,
                              int64_t ndocs, int32_t* ids, int32_t* counts,
                              int max_ids, int unk_id, int threads);

/* Full-size parity without an [ndocs][max_ids] matrix: FNV-1a-64 (SURVEY 8c recipe) of every document's ids. */
int64_t bfo_text_to_ids_digests(const bfo_model* m, const char* utf8, const int64_t* offsets, int64_t ndocs,
                                uint64_t* digests, int32_t* counts, int max_ids, int unk_id, int threads);
void bfo_csr_digests(const void* ids, int elem_size, const int64_t* id_offsets, int64_t ndocs, uint64_t* digests, int threads);
uint64_t bfo_fold_digests(const uint64_t* digests, const int64_t* id_offsets, const int32_t* counts, int64_t ndocs);

#ifdef __cplusplus
}
#endif
#endif
