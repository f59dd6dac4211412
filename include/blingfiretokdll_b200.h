/*
 * blingfiretokdll_b200.h -- C ABI of the B200-native drop-in for BlingFire's TextToIds path.
 *
 * The first block re-declares, with identical names, argument meaning, ownership and return
 * conventions, the entry points of the reference's blingfiretokdll that lie on the
 * TextToIds / TextToWords hot path.  Each declaration cites the reference interface it
 * replaces (paths relative to the reference checkout).  A caller that binds these symbols
 * by name (dist-pypi/blingfire/__init__.py:229-253 via ctypes, nuget/lib/BlingFireUtils.cs:24-35
 * via P/Invoke) can load this library instead of libblingfiretokdll.so without code changes.
 *
 * The second block is additive: the reference has no batch entry point (every export takes
 * one document), and one call per document cannot feed a GPU.
 *
 * (The reference spells the returns `const int`; the qualifier is meaningless on a return
 * type and is dropped here -- the ABI is identical.)
 *
 * Plain C: pointers and sizes only, caller-owned buffers, no exceptions cross this boundary
 * (the reference may throw std::runtime_error through its C ABI, e.g. on a missing model
 * file; here such cases return NULL / 0 / -1).
 *
 * There is NO CPU implementation behind these symbols: every tokenizing call runs on the
 * current CUDA device and fails (0 / -1 / NULL, see BlingFireB200LastError) if no device or
 * kernel image is available.
 */
#ifndef BLINGFIRETOKDLL_B200_H
#define BLINGFIRETOKDLL_B200_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------------------------------
 * Drop-in symbols (same names as the reference)
 * ------------------------------------------------------------------------------------- */

/* blingfiretokdll.h:25, blingfiretokdll.cpp:107-111.  Returns 18000 (algo version 18.0). */
int GetBlingFireTokVersion(void);

/* blingfiretokdll.h:48, blingfiretokdll.cpp:1077-1094.  Loads a compiled .bin LDB, flattens
 * its automata and uploads them to the current CUDA device.  Returns an opaque handle,
 * NULL on any failure (missing file, malformed image, CRC mismatch, unsupported model).
 * The handle also owns run-time state the reference does not have: a device-resident table of words (WordPiece models)
 * or segments (byte-BPE models) the kernels have already resolved, which they extend as they meet new ones.  It is a memo
 * of the reference's own algorithm: results never depend on its content, only the speed does (the first pass over new
 * text is slower than the following ones); it is bounded (64-128 MB per handle) and freed by FreeModel. */
void* LoadModel(const char* pszLdbFileName);

/* blingfiretokdll.h:47, blingfiretokdll.cpp:1055-1071.  Same, from a memory image (copied). */
void* SetModel(const unsigned char* pImgBytes, int ModelByteCount);

/* blingfiretokdll.h:101, blingfiretokdll.cpp:1653-1662.  0 if NULL, else frees and returns 1. */
int FreeModel(void* ModelPtr);

/* blingfiretokdll.h:93-100, blingfiretokdll.cpp:1619-1646.  Writes at most MaxIdsArrLength
 * ids, leaves the rest of pIdsArr untouched, returns the number written.  Returns 0 for a
 * NULL model/text, InUtf8StrByteCount <= 0 or > 1e9, invalid UTF-8 anywhere in the input,
 * or normalisation overflow -- exactly like the reference.  UnkId may be any int. */
int TextToIds(void* ModelPtr, const char* pInUtf8Str, int InUtf8StrByteCount,
                    int32_t* pIdsArr, const int MaxIdsArrLength, const int UnkId);

/* blingfiretokdll.h:57-64 / :75-82, blingfiretokdll.cpp:1320-1335 / :1541-1552. */
int TextToIds_wp(void* ModelPtr, const char* pInUtf8Str, int InUtf8StrByteCount,
                       int32_t* pIdsArr, const int MaxIdsArrLength, const int UnkId);
int TextToIds_sp(void* ModelPtr, const char* pInUtf8Str, int InUtf8StrByteCount,
                       int32_t* pIdsArr, const int MaxIdsArrLength, const int UnkId);

/* blingfiretokdll.h:83-92 / :49-58 / :65-74, blingfiretokdll.cpp:1563-1609 / :1108-1314 / :1349-1535.
 * Same as TextToIds plus, for every id, the byte offsets of the first byte of its first character and
 * of the LAST byte of its last character in the input (UnkId tokens take the offsets of their word).
 * Entries beyond the returned count stay untouched in all three arrays.  For [pos-dict] models a token
 * that consists of the dummy prefix alone has start -1; its end is pinned to -1 (the reference reads
 * the byte before the input there, blingfiretokdll.cpp:1527). */
int TextToIdsWithOffsets(void* ModelPtr, const char* pInUtf8Str, int InUtf8StrByteCount,
                         int32_t* pIdsArr, int* pStartOffsets, int* pEndOffsets,
                         const int MaxIdsArrLength, const int UnkId);
int TextToIdsWithOffsets_wp(void* ModelPtr, const char* pInUtf8Str, int InUtf8StrByteCount,
                            int32_t* pIdsArr, int* pStartOffsets, int* pEndOffsets,
                            const int MaxIdsArrLength, const int UnkId);
int TextToIdsWithOffsets_sp(void* ModelPtr, const char* pInUtf8Str, int InUtf8StrByteCount,
                            int32_t* pIdsArr, int* pStartOffsets, int* pEndOffsets,
                            const int MaxIdsArrLength, const int UnkId);

/* blingfiretokdll.cpp:1669-1679 (exported through blingfiretokdll.def; dist-pypi/blingfire/__init__.py:287-288).
 * Turns the dummy U+2581 prefix of a [pos-dict] model off (true) or on (false) for all later calls on the
 * handle.  Like the reference's, the switch is a plain store: do not flip it while other threads tokenize.
 * Returns 0 for a NULL handle, else 1. */
int SetNoDummyPrefix(void* ModelPtr, bool fNoDummyPrefix);

/* blingfiretokdll.cpp:1689-1745 (dist-pypi/blingfire/__init__.py:256-270).  Concatenates the texts of the ids
 * (model with an [i2w] section, or a separate *.i2w file loaded with LoadModel); with SkipSpecialTokens ids
 * outside the model's regular range are dropped; a space leading the output is dropped.  Returns the length
 * including the trailing NUL (the minimum buffer size that holds the whole output); 0 for a NULL handle, no
 * ids, a model without [i2w] or an id outside the array.  If the return value exceeds
 * MaxOutUtf8StrByteCount the buffer content is undefined, as in the reference.  A table read on the host. */
int IdsToText(void* ModelPtr, const int32_t* pIdsArr, const int IdsCount, char* pOutUtf8Str,
              const int MaxOutUtf8StrByteCount, bool SkipSpecialTokens);

/* blingfiretokdll.h:41, blingfiretokdll.cpp:610-614.  Default word breaker.  Returns -1 on
 * error, 0 for empty input, else the required output size including the trailing NUL (the
 * output is copied only if it fits).  The reference embeds wbd.bin as a byte array; this
 * library loads it from $BLINGFIRE_B200_WBD or <library dir>/wbd.bin on first use. */
int TextToWords(const char* pInUtf8Str, int InUtf8StrByteCount,
                      char* pOutUtf8Str, const int MaxOutUtf8StrByteCount);

/* blingfiretokdll.h:39-40, blingfiretokdll.cpp:597-603.  hModel == NULL selects the default. */
int TextToWordsWithModel(const char* pInUtf8Str, int InUtf8StrByteCount,
                               char* pOutUtf8Str, const int MaxOutUtf8StrByteCount, void* hModel);

/* blingfiretokdll.h:37-38 / :35-36, blingfiretokdll.cpp:415-566 / :575-581.  TextToWords plus, for word
 * k, the byte offset of the first byte of its first character and of the LAST byte of its last
 * character.  Both arrays (either may be NULL) must hold MaxOutUtf8StrByteCount ints and are
 * zero-filled first, like the reference does. */
int TextToWordsWithOffsetsWithModel(const char* pInUtf8Str, int InUtf8StrByteCount, char* pOutUtf8Str,
                                    int* pStartOffsets, int* pEndOffsets, const int MaxOutUtf8StrByteCount, void* hModel);
int TextToWordsWithOffsets(const char* pInUtf8Str, int InUtf8StrByteCount, char* pOutUtf8Str,
                           int* pStartOffsets, int* pEndOffsets, const int MaxOutUtf8StrByteCount);

/* ADDITIVE.  TextToIdsWithOffsets (blingfiretokdll.cpp:1563-1609, -> _wp :1108-1314 / _sp :1349-1535 with offsets) for a
 * batch: documents as CSR; pIds / pStarts / pEnds are row-major [DocCount][MaxIdsPerDoc], pCounts[i] = what the per-document
 * call returns for document i; entries of a row beyond its count stay untouched.  Byte offsets count from the document's
 * first byte, with the reference's conventions (end = last byte of the last character; -1 for the dummy prefix).
 * Returns the total number of ids, -1 on error. */
int64_t TextToIdsWithOffsetsBatch(void* ModelPtr, const char* pUtf8, const int64_t* pOffsets, int64_t DocCount, int32_t* pIds,
                                  int32_t* pStarts, int32_t* pEnds, int32_t* pCounts, int MaxIdsPerDoc, int UnkId);

/* ADDITIVE.  The same with compact output, like TextToIdsBatchCsr: ids / starts / ends of document i are entries
 * [pIdOffsets[i] .. pIdOffsets[i+1]) of the three arrays (CsrCapacity entries each; 12 bytes per id cross PCIe instead of
 * 12 * MaxIdsPerDoc per document).  pIdOffsets has DocCount+1 entries and is always complete on a non-error return.
 * Returns the total number of ids; if it exceeds CsrCapacity, negated (the arrays then hold the chunks that still
 * fitted); -1 on error. */
int64_t TextToIdsWithOffsetsBatchCsr(void* ModelPtr, const char* pUtf8, const int64_t* pOffsets, int64_t DocCount, int32_t* pIdsCsr,
                                     int32_t* pStartsCsr, int32_t* pEndsCsr, int64_t CsrCapacity, int64_t* pIdOffsets,
                                     int MaxIdsPerDoc, int UnkId);

/* ADDITIVE (not in the reference, which takes one document per call: blingfiretokdll.cpp:415-566).  TextToWords[WithModel]
 * for a batch: documents as CSR (pUtf8, pOffsets[DocCount+1]), strings as CSR.  The lexer and the string building
 * (:507-555) both run on the GPU.  hModel may be NULL (the default word breaker, like TextToWords).
 *   pResults[i]     what TextToWordsWithModel would return for document i: -1 (bad UTF-8), 0 (empty input), else the byte
 *                   length of its string INCLUDING the trailing NUL
 *   pOutOffsets[i]  where that string starts in pOut; documents without one take no bytes; pOutOffsets[DocCount] = total
 * Returns the total bytes; -total when Capacity is too small (pOutOffsets / pResults are complete, call again); -1 on
 * error (BlingFireB200LastError). */
int64_t TextToWordsBatch(void* hModel, const char* pUtf8, const int64_t* pOffsets, int64_t DocCount, char* pOut,
                         int64_t Capacity, int64_t* pOutOffsets, int32_t* pResults);
/* ADDITIVE.  The same for TextToSentences[WithModel] (blingfiretokdll.cpp:163-355): one string per document, its sentences
 * joined by '\n'; hModel NULL = the default sentence breaker. */
int64_t TextToSentencesBatch(void* hModel, const char* pUtf8, const int64_t* pOffsets, int64_t DocCount, char* pOut,
                             int64_t Capacity, int64_t* pOutOffsets, int32_t* pResults);

/* blingfiretokdll.h:33 / :31-32 / :29-30 / :27-28, blingfiretokdll.cpp:398-401 / :378-381 / :364-368 / :163-355.
 * Splits a paragraph into sentences, '\n'-joined (a '\n' inside a sentence becomes ' '); same return
 * convention as TextToWords.  The sentence-breaking lexer (default: sbd.bin, loaded from
 * $BLINGFIRE_B200_SBD or <library dir>/sbd.bin on first use) runs on the GPU. */
int TextToSentences(const char* pInUtf8Str, int InUtf8StrByteCount, char* pOutUtf8Str, const int MaxOutUtf8StrByteCount);
int TextToSentencesWithModel(const char* pInUtf8Str, int InUtf8StrByteCount, char* pOutUtf8Str,
                             const int MaxOutUtf8StrByteCount, void* hModel);
int TextToSentencesWithOffsets(const char* pInUtf8Str, int InUtf8StrByteCount, char* pOutUtf8Str,
                               int* pStartOffsets, int* pEndOffsets, const int MaxOutUtf8StrByteCount);
int TextToSentencesWithOffsetsWithModel(const char* pInUtf8Str, int InUtf8StrByteCount, char* pOutUtf8Str,
                                        int* pStartOffsets, int* pEndOffsets, const int MaxOutUtf8StrByteCount, void* hModel);

/* ---------------------------------------------------------------------------------------
 * Additive batch entry points (new)
 * ------------------------------------------------------------------------------------- */

/* Documents are a CSR byte buffer: document i = pUtf8[pOffsets[i] .. pOffsets[i+1]).
 * For every i the pair (pCounts[i], pIds[i*MaxIdsPerDoc .. +pCounts[i])) equals what the
 * reference's TextToIds returns for document i with the same MaxIdsArrLength and UnkId;
 * entries of a row beyond pCounts[i] are left untouched.  HOST pointers; host<->device
 * copies are pipelined inside the call.  Returns the total number of ids, or -1 on error. */
int64_t TextToIdsBatch(void* ModelPtr, const char* pUtf8, const int64_t* pOffsets, int64_t DocCount,
                       int32_t* pIds, int32_t* pCounts, int MaxIdsPerDoc, int UnkId);

/* Same contract, compact output: ids of document i are pIdsCsr[pIdOffsets[i] .. pIdOffsets[i+1]),
 * pIdOffsets has DocCount+1 entries (always complete on a non-error return).  pIdsCsr must hold CsrCapacity
 * ids (a document of n bytes yields at most n ids for lexer models, n+1 for [pos-dict] models, 2n+4 for
 * [pos-dict] models with a charmap; never more than MaxIdsPerDoc).  If the batch produces more, the required
 * capacity is returned negated and the content of pIdsCsr is undefined (the chunks that still fitted have
 * been copied).  HOST pointers, pinned or pageable: pageable buffers are staged through library-owned
 * pinned memory by a few copy threads on the GPU's NUMA node.  This is the call bench.py times end to end. */
int64_t TextToIdsBatchCsr(void* ModelPtr, const char* pUtf8, const int64_t* pOffsets, int64_t DocCount,
                          int32_t* pIdsCsr, int64_t CsrCapacity, int64_t* pIdOffsets,
                          int MaxIdsPerDoc, int UnkId);

/* Same with 16-bit ids, for models whose ids all fit (every shipped WordPiece/BPE model below 65 536 entries)
 * and 0 <= UnkId <= 65535: halves the device->host bytes.  -1 if the model or UnkId does not qualify. */
int64_t TextToIdsBatchCsrU16(void* ModelPtr, const char* pUtf8, const int64_t* pOffsets, int64_t DocCount,
                             uint16_t* pIdsCsr, int64_t CsrCapacity, int64_t* pIdOffsets,
                             int MaxIdsPerDoc, int UnkId);

/* DEVICE pointers on the model's device, no copies, asynchronous on `cudaStream` (a
 * cudaStream_t, NULL = legacy default stream).  dUtf8 must be 4-byte aligned with at least
 * 8 readable bytes of slack after TotalBytes (any cudaMalloc'ed buffer qualifies); for the generic-lexer
 * engine dOffsets[0] must be 0.  dIds is [DocCount][MaxIdsPerDoc] row-major.  Calls on different streams may
 * overlap (each call has its own work counter; [pos-dict] scratch is kept per stream).  For [pos-dict]
 * models this form reads the longest document back from the device first (one stream synchronisation);
 * TextToIdsBatchDeviceSized takes it from the caller (MaxDocBytes >= every document's byte length) and
 * stays asynchronous.  Returns 0 on success, -1 on error. */
int TextToIdsBatchDevice(void* ModelPtr, const char* dUtf8, const int64_t* dOffsets, int64_t DocCount,
                         int64_t TotalBytes, int32_t* dIds, int32_t* dCounts,
                         int MaxIdsPerDoc, int UnkId, void* cudaStream);
int TextToIdsBatchDeviceSized(void* ModelPtr, const char* dUtf8, const int64_t* dOffsets, int64_t DocCount,
                              int64_t TotalBytes, int64_t MaxDocBytes, int32_t* dIds, int32_t* dCounts,
                              int MaxIdsPerDoc, int UnkId, void* cudaStream);

/* After device-pointer calls on `cudaStream`: synchronises the stream and returns 0, or the error code a
 * [pos-dict] kernel raised because its scratch could not hold a document (it never guesses); -1 on a CUDA error. */
int BlingFireB200DeviceStatus(void* ModelPtr, void* cudaStream);

/* Row-major device ids -> CSR on the device: dRowOff[DocCount+1] = exclusive prefix sum of dCounts,
 * dCsr[dRowOff[i] + k] = dIds[i*MaxIdsPerDoc + k] for k < dCounts[i].  Asynchronous.  0 / -1. */
int64_t BlingFireB200CompactDevice(const int32_t* dIds, const int32_t* dCounts, int64_t DocCount, int MaxIdsPerDoc,
                                   int32_t* dCsr, int64_t* dRowOff, void* cudaStream);

/* Last error of the calling thread ("" if none).  Never NULL. */
const char* BlingFireB200LastError(void);

/* Number of GPU kernels this library has launched in this process (bench.py: gpu_launches). */
int64_t BlingFireB200KernelLaunches(void);

/* Device time (ms, CUDA events) the tokenization kernels of the calling thread's last host batch call
 * took, summed over its chunks -- the kernel share of an end-to-end call. */
double BlingFireB200LastKernelMs(void);

/* Which engine serves the model: 1 = fused WordPiece kernel (FastPath lexer models),
 * 2 = generic lexer engine, 3 = segmentation engine (Unigram-LM / BPE), 0 = none. */
int BlingFireB200ModelEngine(void* ModelPtr);

#ifdef __cplusplus
}
#endif
#endif
