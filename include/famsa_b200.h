/*
 * famsa_b200.h -- C ABI of libfamsa_b200.so, the B200 (sm_100a) replacement for FAMSA's two
 * data-parallel hot paths.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the FAMSA
 * source tree, v2.5.0).  INTEGRATION.md shows the C++ shims a FAMSA maintainer would add so that
 * src/tree/ *, src/msa.cpp and src/msa_refinement.cpp compile and run unchanged on top of this.
 *
 * Conventions
 *   - All functions return 0 on success and a non-zero FAMSA_E_* code on failure;
 *     famsa_last_error() returns a human-readable message for the calling thread's last failure
 *     (the reference has no status codes: errors are std::runtime_error caught in main(),
 *     src/famsa.cpp:160-165 -- the C++ shim rethrows the message).
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails.
 *   - A context is bound to one device.  Calls on one context are serialised internally, so the
 *     reference's concurrent per-thread CLCSBP instances may share one context.
 *   - Residue codes are the reference's: 0..23 = "ARNDCQEGHILKMFPSTWYVBZX*" (src/core/sequence.cpp:17),
 *     22 = unknown / padding.  Only codes < 20 can match (sequence.cpp:199).
 */
#ifndef FAMSA_B200_H
#define FAMSA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FAMSA_B200_ABI_VERSION 1

enum {
    FAMSA_OK = 0,
    FAMSA_E_INVALID = 1,   /* bad argument */
    FAMSA_E_NO_DEVICE = 2, /* no usable CUDA device (there is no CPU fallback) */
    FAMSA_E_CUDA = 3,      /* CUDA runtime / kernel failure */
    FAMSA_E_STATE = 4,     /* call order violated (e.g. LCS before upload) */
    FAMSA_E_NOMEM = 5
};

typedef struct famsa_ctx famsa_ctx;

/* ------------------------------------------------------------------ context */

int famsa_abi_version(void);

/* device = CUDA ordinal, or -1 for the current device. */
int famsa_create(int device, famsa_ctx** out_ctx);
void famsa_destroy(famsa_ctx* ctx);

/* Message for the last failure on the calling thread ("" if none). Never NULL. */
const char* famsa_last_error(void);

/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t famsa_kernel_launches(const famsa_ctx* ctx);

/* ------------------------------------------------------------------ HP-1: all-pairs LCS length
 *
 * Replaces, per worker thread of every guide-tree builder:
 *   CSequence::ComputeBitMasks                 src/core/sequence.cpp:190-201
 *   MSTPrim::prepare_sequences_view/bit_masks  src/tree/MSTPrim.cpp:554-566, 837-852
 *   CLCSBP::GetLCSBP (4 overloads)             src/lcs/lcsbp.h:38-46, lcsbp.cpp:48-368
 *   CLCSBP_{Classic,AVX,AVX2,AVX512,NEON}      src/lcs/lcsbp_classic.h, src/simd/lcsbp_*_intr.h
 * behind the four batch drivers
 *   AbstractTreeGenerator::calculateDistanceVector / Range / RangeSV / Matrix
 *                                              src/tree/AbstractTreeGenerator.hpp:131,191,288,379
 * Results are the raw LCS lengths (what GetLCSBP writes to `dist`); the LCS -> distance
 * Transform (AbstractTreeGenerator.hpp:28-82) stays on the host so distances are bit-identical.
 * Lengths are bit-exact against the reference, including its dropped-carry corner
 * (lcsbp_classic.h:55-56): which sequence is the row (`seq0`, supplies the bit masks) matters.
 */

/* Upload a sequence set.  Sequence i is codes[offsets[i] .. offsets[i]+lens[i]) (no padding
 * needed).  Replaces any previously uploaded set.  Builds the device-side bit-mask tables. */
int famsa_lcs_upload(famsa_ctx* ctx, const int8_t* codes, const uint64_t* offsets,
                     const uint32_t* lens, uint32_t n_seqs);

uint32_t famsa_lcs_n_seqs(const famsa_ctx* ctx);

/* Lower triangle, rows [row_begin, row_end): for each row i and each j < i the LCS length with
 * sequence i as the row (seq0).  Layout = TriangleMatrix::access (src/tree/TreeDefs.h:114-119):
 *   out[ i(i-1)/2 - row_begin(row_begin-1)/2 + j ].
 * elem_bytes is 2 (uint16_t, requires every length < 65536) or 4 (uint32_t, the type GetLCSBP
 * writes).  `out` is HOST memory in famsa_lcs_triangle and DEVICE memory in the _device variant;
 * `stream` is a cudaStream_t (NULL = the context's own stream; the call then also synchronises). */
int famsa_lcs_triangle(famsa_ctx* ctx, uint32_t row_begin, uint32_t row_end, void* out,
                       int elem_bytes);
int famsa_lcs_triangle_device(famsa_ctx* ctx, uint32_t row_begin, uint32_t row_end, void* d_out,
                              int elem_bytes, void* stream);

/* n_ref reference rows against a column list: out[r * n_col + k] = LCS length with sequence
 * ref_ids[r] as the row (seq0) and sequence col_ids[k] streamed.  col_ids == NULL means columns
 * 0..n_col-1 (calculateDistanceVector's prefix shape).  This is calculateDistanceRange /
 * RangeSV (Prim vertex vs unvisited set, medoid seed vs members) with any number of rows. */
int famsa_lcs_rows(famsa_ctx* ctx, const uint32_t* ref_ids, uint32_t n_ref,
                   const uint32_t* col_ids, uint32_t n_col, void* out, int elem_bytes);
int famsa_lcs_rows_device(famsa_ctx* ctx, const uint32_t* d_ref_ids, uint32_t n_ref,
                          const uint32_t* d_col_ids, uint32_t n_col, void* d_out, int elem_bytes,
                          void* stream);

/* Host-side Transform<T, Distance> (AbstractTreeGenerator.hpp:28-82), provided so bindings that
 * are not C++ get bit-identical distances.  kind: 0 indel075_div_lcs, 1 indel_div_lcs,
 * 2 pairwise_identity. */
double famsa_transform_f64(int kind, uint32_t lcs, uint32_t len1, uint32_t len2);
float famsa_transform_f32(int kind, uint32_t lcs, uint32_t len1, uint32_t len2);

/* Timing of the most recent LCS call on this context, measured with CUDA events on the stream the
 * kernels ran on: total = all kernels of the call, main = the lcs tile kernel(s) only. */
int famsa_lcs_last_timing(const famsa_ctx* ctx, float* total_ms, float* main_kernel_ms,
                          uint64_t* n_pairs);

#ifdef __cplusplus
}
#endif
#endif /* FAMSA_B200_H */
