// C ABI of libfamsa_b200.so -- see include/famsa_b200.h for the contract of every entry point.
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <numeric>
#include <algorithm>
#include <vector>

#include "ctx.h"

namespace fb {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const char* get_error() { return g_error.c_str(); }

int scratch_acquire(famsa_ctx* ctx, cudaStream_t st)
{
    if (ctx->busy) FB_CUDA(cudaStreamWaitEvent(st, ctx->ev_busy, 0));
    return FAMSA_OK;
}
int scratch_release(famsa_ctx* ctx, cudaStream_t st, bool synced)
{
    if (synced) { ctx->busy = false; return FAMSA_OK; }
    FB_CUDA(cudaEventRecord(ctx->ev_busy, st));
    ctx->busy = true;
    return FAMSA_OK;
}
} // namespace fb

using fb::set_error;

#define FB_CHECK_CTX(ctx)                          \
    if (!(ctx)) {                                  \
        set_error("famsa_ctx is NULL");            \
        return FAMSA_E_INVALID;                    \
    }

extern "C" {

int famsa_abi_version(void) { return FAMSA_B200_ABI_VERSION; }

const char* famsa_last_error(void) { return fb::get_error(); }

int famsa_create(int device, famsa_ctx** out_ctx)
{
    if (!out_ctx) { set_error("out_ctx is NULL"); return FAMSA_E_INVALID; }
    *out_ctx = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                  "); libfamsa_b200 has no CPU fallback");
        return FAMSA_E_NO_DEVICE;
    }
    if (device < 0) FB_CUDA(cudaGetDevice(&device));
    if (device >= count) { set_error("device ordinal out of range"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop{};
    FB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                  std::to_string(prop.minor) + "; this library is built for sm_100a only");
        return FAMSA_E_NO_DEVICE;
    }
    famsa_ctx* ctx = new (std::nothrow) famsa_ctx();
    if (!ctx) { set_error("out of host memory"); return FAMSA_E_NOMEM; }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    const int rc = [&]() -> int {
        FB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        for (auto& ev : ctx->ev) FB_CUDA(cudaEventCreate(&ev));
        FB_CUDA(cudaEventCreateWithFlags(&ctx->ev_busy, cudaEventDisableTiming));
        FB_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        for (auto& e : ctx->ev_copy) FB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : ctx->ev_join) FB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& a : ctx->aux_stream) FB_CUDA(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
        // per-call scratch and resident profiles are stream-ordered allocations: the pool keeps what it has
        cudaMemPool_t pool;
        FB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        unsigned long long keep = ~0ull;
        FB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        return FAMSA_OK;
    }();
    if (rc) { famsa_destroy(ctx); return rc; }
    *out_ctx = ctx;
    return FAMSA_OK;
}

void famsa_destroy(famsa_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    fb::LcsState& S = ctx->lcs;
    for (fb::DevBuf* b : {&S.d_perm, &S.d_invperm, &S.d_len_sorted, &S.d_code_off, &S.d_codes, &S.d_blob,
                          &S.d_group_blob, &S.d_raw_codes, &S.d_raw_off, &S.d_raw_len, &S.d_flags, &S.d_pow075, &S.d_assign_lcs, &S.d_pow075_f64, &S.d_prim_tri, &S.d_prim_side, &S.d_prim_state,
                          &S.d_prim_out, &S.d_prim_sideidx, &S.d_prim_cand, &S.d_prim_dtri, &S.d_prim_comp, &S.d_prim_best, &S.d_prim_part,
                          &S.d_assign, &S.d_mind, &S.d_tiles,
                          &S.d_res, &S.d_refpos, &S.d_ids_a, &S.d_ids_b, &S.d_out_stage, &S.d_masks64, &S.d_x64})
        b->release();
    fb::prof_release_all(ctx);
    fb::DpState& D = ctx->dp;
    if (D.h_pinned) cudaFreeHost(D.h_pinned);
    for (fb::DevBuf* b : {&D.d_dirs_out, &D.d_tables, &D.d_results, &D.d_path})
        b->release();
    for (auto& ev : ctx->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_block)
        if (ev) cudaEventDestroy(ev);
    for (auto& ev : ctx->ev_host)
        if (ev) cudaEventDestroy(ev);
    if (ctx->ev_busy) cudaEventDestroy(ctx->ev_busy);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    for (auto& e : ctx->ev_copy)
        if (e) cudaEventDestroy(e);
    for (auto& st : ctx->peer_stream)
        if (st) cudaStreamDestroy(st);
    for (auto& e : ctx->ev_join)
        if (e) cudaEventDestroy(e);
    for (auto& a : ctx->aux_stream)
        if (a) cudaStreamDestroy(a);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

uint64_t famsa_kernel_launches(const famsa_ctx* ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------ HP-1

int famsa_lcs_upload(famsa_ctx* ctx, const int8_t* codes, const uint64_t* offsets, const uint32_t* lens,
                     uint32_t n_seqs)
{
    FB_CHECK_CTX(ctx);
    if (n_seqs && (!codes || !offsets || !lens)) { set_error("NULL sequence arrays"); return FAMSA_E_INVALID; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = fb::scratch_acquire(ctx, ctx->stream);                 // earlier *_device calls may still read the old set
    if (rc) return rc;
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    fb::scratch_release(ctx, ctx->stream, true);
    ctx->lcs.h_order.clear();
    return fb::lcs_upload(ctx, codes, offsets, lens, n_seqs);
}

int famsa_lcs_upload_sorted(famsa_ctx* ctx, const int8_t* codes, const uint64_t* offsets, const uint32_t* lens, uint32_t n_seqs)
{
    FB_CHECK_CTX(ctx);
    if (n_seqs && (!codes || !offsets || !lens)) { set_error("NULL sequence arrays"); return FAMSA_E_INVALID; }
    // the permutation the library would apply internally (length descending, ties in caller order), applied up front: every
    // index of the calls that follow is then a position in that order
    std::vector<uint32_t> order(n_seqs);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return lens[x] > lens[y]; });
    std::vector<uint64_t> off(n_seqs);
    std::vector<uint32_t> len(n_seqs);
    uint64_t total = 0;
    for (uint32_t p = 0; p < n_seqs; ++p) { off[p] = total; len[p] = lens[order[p]]; total += len[p]; }
    std::vector<int8_t> packed(std::max<uint64_t>(total, 1));
    for (uint32_t p = 0; p < n_seqs; ++p) memcpy(packed.data() + off[p], codes + offsets[order[p]], len[p]);
    const int rc = famsa_lcs_upload(ctx, packed.data(), off.data(), len.data(), n_seqs);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->lcs.h_order = std::move(order);
    return FAMSA_OK;
}

int famsa_lcs_sorted_order(famsa_ctx* ctx, uint32_t* sorted_to_caller)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!sorted_to_caller && ctx->lcs.n) { set_error("sorted_to_caller is NULL"); return FAMSA_E_INVALID; }
    if (ctx->lcs.h_order.size() == ctx->lcs.n) std::copy(ctx->lcs.h_order.begin(), ctx->lcs.h_order.end(), sorted_to_caller);
    else std::iota(sorted_to_caller, sorted_to_caller + ctx->lcs.n, 0u);          // plain famsa_lcs_upload: indices are the caller's
    return FAMSA_OK;
}

uint64_t famsa_lcs_last_tiles(const famsa_ctx* ctx) { return ctx ? ctx->lcs.last_tiles : 0; }

uint32_t famsa_lcs_n_seqs(const famsa_ctx* ctx) { return ctx ? ctx->lcs.n : 0; }

static int check_elem(const famsa_ctx* ctx, int elem_bytes)
{
    if (elem_bytes != 2 && elem_bytes != 4) { set_error("elem_bytes must be 2 or 4"); return FAMSA_E_INVALID; }
    if (elem_bytes == 2 && ctx->lcs.max_len >= 65536) {
        set_error("elem_bytes == 2 needs every sequence shorter than 65536");
        return FAMSA_E_INVALID;
    }
    return FAMSA_OK;
}

static int finish_timing(famsa_ctx* ctx)
{
    float total = 0.f, main_ms = 0.f;
    FB_CUDA(cudaEventSynchronize(ctx->ev[3]));
    FB_CUDA(cudaEventElapsedTime(&total, ctx->ev[0], ctx->ev[3]));
    FB_CUDA(cudaEventElapsedTime(&main_ms, ctx->ev[1], ctx->ev[2]));
    ctx->lcs.last_total_ms = total;
    ctx->lcs.last_main_ms = main_ms;
    return FAMSA_OK;
}

static int triangle_locked(famsa_ctx* ctx, uint32_t row_begin, uint32_t row_end, void* d_out, int elem_bytes,
                           void* stream)
{
    if (ctx->lcs.n == 0 && row_end > 0) { set_error("famsa_lcs_upload has not been called"); return FAMSA_E_STATE; }
    if (row_begin > row_end || row_end > ctx->lcs.n) { set_error("row range out of bounds"); return FAMSA_E_INVALID; }
    int rc = check_elem(ctx, elem_bytes);
    if (rc) return rc;
    if (!d_out && row_end > row_begin && row_end > 1) { set_error("d_out is NULL"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    if ((rc = fb::scratch_acquire(ctx, st))) return rc;
    rc = fb::lcs_triangle(ctx, row_begin, row_end, d_out, elem_bytes, st);
    if (rc) return rc;
    if (!stream) { FB_CUDA(cudaStreamSynchronize(st)); fb::scratch_release(ctx, st, true); return finish_timing(ctx); }
    return fb::scratch_release(ctx, st, false);
}

int famsa_lcs_triangle_device(famsa_ctx* ctx, uint32_t row_begin, uint32_t row_end, void* d_out, int elem_bytes,
                              void* stream)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    return triangle_locked(ctx, row_begin, row_end, d_out, elem_bytes, stream);
}

int famsa_lcs_triangle(famsa_ctx* ctx, uint32_t row_begin, uint32_t row_end, void* out, int elem_bytes)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (row_begin > row_end || row_end > ctx->lcs.n) { set_error("row range out of bounds"); return FAMSA_E_INVALID; }
    if (elem_bytes != 2 && elem_bytes != 4) { set_error("elem_bytes must be 2 or 4"); return FAMSA_E_INVALID; }
    auto tri = [](uint64_t r) { return r ? r * (r - 1) / 2 : 0; };
    const uint64_t pairs = tri(row_end) - tri(row_begin);
    if (pairs && !out) { set_error("out is NULL"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = ctx->lcs.d_out_stage.reserve(std::max<uint64_t>(pairs, 1) * elem_bytes);
    if (rc) return rc;
    char* d_out = static_cast<char*>(ctx->lcs.d_out_stage.p);

    // Row blocks with equal numbers of pairs: all kernels are queued first, then every block is copied back as
    // soon as its event fires, so the D2H of block k overlaps the kernels of blocks k+1...
    constexpr int kBlocks = 8;
    // (worth it only when the copy is long compared with the tail of a kernel: hundreds of millions of pairs)
    unsigned long long block_min = 400000000ull;
    if (const char* e = getenv("FAMSA_LCS_BLOCK_MIN_PAIRS")) block_min = strtoull(e, nullptr, 10);   // test knob
    const int n_blocks = (ctx->lcs.identity_perm && pairs > block_min) ? kBlocks : 1;
    uint32_t bounds[kBlocks + 1];
    bounds[0] = row_begin;
    for (int b = 1; b < n_blocks; ++b) {
        const double target = (double)tri(row_begin) + (double)pairs * b / n_blocks;
        uint32_t r = (uint32_t)((1.0 + std::sqrt(1.0 + 8.0 * target)) / 2.0);
        r = (r + 16) / 32 * 32;                                        // on a mask-group boundary: no group is computed twice
        bounds[b] = std::min(std::max(r, bounds[b - 1]), row_end);
    }
    bounds[n_blocks] = row_end;
    if (!ctx->copy_stream) {
        FB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (auto& e : ctx->ev_block) FB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : ctx->ev_host) FB_CUDA(cudaEventCreate(&e));
    }
    if (ctx->lcs.n == 0 && row_end > 0) { set_error("famsa_lcs_upload has not been called"); return FAMSA_E_STATE; }
    rc = check_elem(ctx, elem_bytes);
    if (rc) return rc;
    if ((rc = fb::scratch_acquire(ctx, ctx->stream))) return rc;
    FB_CUDA(cudaEventRecord(ctx->ev_host[0], ctx->stream));
    float main_ms = 0.f;
    rc = fb::lcs_triangle(ctx, row_begin, row_end, d_out, elem_bytes, ctx->stream, bounds, n_blocks, ctx->ev_block, true, true);
    if (rc) return rc;
    if (!ctx->lcs.identity_perm && n_blocks == 1) FB_CUDA(cudaEventRecord(ctx->ev_block[0], ctx->stream));
    FB_CUDA(cudaEventRecord(ctx->ev_host[1], ctx->stream));
    for (int b = 0; b < n_blocks; ++b) {
        const uint64_t off = (tri(bounds[b]) - tri(row_begin)) * elem_bytes;
        const uint64_t bytes = (tri(bounds[b + 1]) - tri(bounds[b])) * elem_bytes;
        if (!bytes) continue;
        FB_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_block[b], 0));
        FB_CUDA(cudaMemcpyAsync(static_cast<char*>(out) + off, d_out + off, bytes, cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    FB_CUDA(cudaStreamSynchronize(ctx->copy_stream));
    fb::scratch_release(ctx, ctx->stream, true);
    FB_CUDA(cudaEventElapsedTime(&main_ms, ctx->ev_host[0], ctx->ev_host[1]));
    ctx->lcs.last_total_ms = main_ms;
    ctx->lcs.last_main_ms = main_ms;
    ctx->lcs.last_pairs = pairs;
    return FAMSA_OK;
}

// ------------------------------------------------------------------ multi-GPU exchange over peer memory

int famsa_device_alloc(famsa_ctx* ctx, uint64_t bytes, void** d_ptr)
{
    FB_CHECK_CTX(ctx);
    if (!d_ptr) { set_error("d_ptr is NULL"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(ctx->device));
    *d_ptr = nullptr;
    const cudaError_t e = cudaMalloc(d_ptr, std::max<uint64_t>(bytes, 1));   // a whole allocation: exportable as it is
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); set_error("out of device memory (" + std::to_string(bytes) + " bytes)"); return FAMSA_E_NOMEM; }
    FB_CUDA(e);
    return FAMSA_OK;
}

int famsa_device_free(famsa_ctx* ctx, void* d_ptr)
{
    FB_CHECK_CTX(ctx);
    FB_CUDA(cudaSetDevice(ctx->device));
    FB_CUDA(cudaDeviceSynchronize());
    FB_CUDA(cudaFree(d_ptr));
    return FAMSA_OK;
}

int famsa_ipc_export(famsa_ctx* ctx, void* d_ptr, uint8_t handle[64])
{
    FB_CHECK_CTX(ctx);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
    if (!d_ptr || !handle) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    FB_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle, &h, 64);
    return FAMSA_OK;
}

int famsa_ipc_open(famsa_ctx* ctx, const uint8_t handle[64], void** d_ptr)
{
    FB_CHECK_CTX(ctx);
    if (!d_ptr || !handle) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    FB_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return FAMSA_OK;
}

int famsa_ipc_close(famsa_ctx* ctx, void* d_ptr)
{
    FB_CHECK_CTX(ctx);
    FB_CUDA(cudaSetDevice(ctx->device));
    FB_CUDA(cudaDeviceSynchronize());
    FB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return FAMSA_OK;
}

int famsa_lcs_triangle_exchange(famsa_ctx* ctx, uint32_t row_begin, uint32_t row_end, void* d_full, void* const* d_peer_full,
                                uint32_t n_peers, int elem_bytes, uint32_t n_pieces, void* stream)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    constexpr uint32_t kPieces = sizeof(ctx->ev_block) / sizeof(ctx->ev_block[0]);
    if (ctx->lcs.n == 0) { set_error("famsa_lcs_upload has not been called"); return FAMSA_E_STATE; }
    if (row_begin > row_end || row_end > ctx->lcs.n) { set_error("row range out of bounds"); return FAMSA_E_INVALID; }
    if (!d_full || (n_peers && !d_peer_full)) { set_error("NULL triangle buffer"); return FAMSA_E_INVALID; }
    for (uint32_t k = 0; k < n_peers; ++k)
        if (!d_peer_full[k]) { set_error("NULL peer buffer"); return FAMSA_E_INVALID; }
    int rc = check_elem(ctx, elem_bytes);
    if (rc) return rc;
    FB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    if (!ctx->copy_stream) {
        FB_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (auto& e : ctx->ev_block) FB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : ctx->ev_host) FB_CUDA(cudaEventCreate(&e));
    }
    auto tri = [](uint64_t r) { return r ? r * (r - 1) / 2 : 0; };
    // Pieces (a non-identity order cannot be cut by rows: one piece).  The copy of the LAST piece is the only exposed part of
    // the exchange, so the pieces shrink towards the end: 14 % of the pairs each at first, 8 % at last.
    const uint32_t np = ctx->lcs.identity_perm ? std::min(std::max(n_pieces, 1u), kPieces) : 1u;
    uint32_t bounds[kPieces + 1];
    bounds[0] = row_begin;
    const uint64_t pairs = tri(row_end) - tri(row_begin);
    static const double kCut8[9] = {0.0, 0.14, 0.28, 0.42, 0.56, 0.70, 0.82, 0.92, 1.0};
    for (uint32_t b = 1; b < np; ++b) {
        const double frac = np == 8 ? kCut8[b] : (double)b / np;
        const double target = (double)tri(row_begin) + (double)pairs * frac;
        uint32_t r = (uint32_t)((1.0 + std::sqrt(1.0 + 8.0 * target)) / 2.0);
        r = (r + 16) / 32 * 32;                                        // on a mask-group boundary: no group is computed twice
        bounds[b] = std::min(std::max(r, bounds[b - 1]), row_end);
    }
    bounds[np] = row_end;
    if ((rc = fb::scratch_acquire(ctx, st))) return rc;
    char* own = static_cast<char*>(d_full) + tri(row_begin) * elem_bytes;    // lcs_triangle indexes from its first row
    rc = fb::lcs_triangle(ctx, row_begin, row_end, own, elem_bytes, st, bounds, (int)np, ctx->ev_block, true, true);
    if (rc) return rc;
    // every finished piece goes to the same place of every peer's triangle while the next pieces are being computed: copy
    // engines over NVLink, no SM and no collective kernel involved
    if (n_peers) {
        constexpr int kLanes = 1 + sizeof(ctx->peer_stream) / sizeof(ctx->peer_stream[0]);
        cudaStream_t lanes[kLanes];
        lanes[0] = ctx->copy_stream;
        for (int a = 1; a < kLanes; ++a) {
            if (!ctx->peer_stream[a - 1]) FB_CUDA(cudaStreamCreateWithFlags(&ctx->peer_stream[a - 1], cudaStreamNonBlocking));
            lanes[a] = ctx->peer_stream[a - 1];
        }
        const int n_lanes = (int)std::min<uint32_t>(n_peers, kLanes);
        for (uint32_t b = 0; b < np; ++b) {
            const uint64_t off = tri(bounds[b]) * elem_bytes, bytes = (tri(bounds[b + 1]) - tri(bounds[b])) * elem_bytes;
            if (!bytes) continue;
            for (int a = 0; a < n_lanes; ++a) FB_CUDA(cudaStreamWaitEvent(lanes[a], ctx->ev_block[b], 0));
            for (uint32_t k = 0; k < n_peers; ++k)
                FB_CUDA(cudaMemcpyAsync(static_cast<char*>(d_peer_full[k]) + off, static_cast<char*>(d_full) + off, bytes,
                                        cudaMemcpyDeviceToDevice, lanes[k % n_lanes]));
        }
        for (int a = 0; a < n_lanes; ++a) {                           // the caller's stream continues after the last copy
            FB_CUDA(cudaEventRecord(ctx->ev_copy[a], lanes[a]));
            FB_CUDA(cudaStreamWaitEvent(st, ctx->ev_copy[a], 0));
        }
    }
    if (!stream) { FB_CUDA(cudaStreamSynchronize(st)); fb::scratch_release(ctx, st, true); return finish_timing(ctx); }
    return fb::scratch_release(ctx, st, false);
}

static int rows_common(famsa_ctx* ctx, const uint32_t* d_ref_ids, const uint32_t* h_ref_ids, uint32_t n_ref,
                       const uint32_t* d_col_ids, uint32_t n_col, void* d_out, int elem_bytes, void* stream)
{
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    for (uint32_t r = 0; r < n_ref; ++r)
        if (h_ref_ids[r] >= ctx->lcs.n) { set_error("ref id out of range"); return FAMSA_E_INVALID; }
    int rc = fb::scratch_acquire(ctx, st);
    if (rc) return rc;
    rc = fb::lcs_rows(ctx, d_ref_ids, h_ref_ids, n_ref, d_col_ids, n_col, d_out, elem_bytes, st);
    if (rc) return rc;
    if (!stream) { FB_CUDA(cudaStreamSynchronize(st)); fb::scratch_release(ctx, st, true); return finish_timing(ctx); }
    return fb::scratch_release(ctx, st, false);
}

int famsa_lcs_rows_device(famsa_ctx* ctx, const uint32_t* d_ref_ids, uint32_t n_ref, const uint32_t* d_col_ids,
                          uint32_t n_col, void* d_out, int elem_bytes, void* stream)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->lcs.n == 0) { set_error("famsa_lcs_upload has not been called"); return FAMSA_E_STATE; }
    int rc = check_elem(ctx, elem_bytes);
    if (rc) return rc;
    if (!d_col_ids && n_col > ctx->lcs.n) { set_error("n_col exceeds the number of sequences"); return FAMSA_E_INVALID; }
    if (n_ref && !d_ref_ids) { set_error("d_ref_ids is NULL"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(ctx->device));
    std::vector<uint32_t> h_ref(n_ref);
    if (n_ref) {      // ordered on the caller's stream (a blocking copy on the legacy stream would not be)
        cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
        FB_CUDA(cudaMemcpyAsync(h_ref.data(), d_ref_ids, sizeof(uint32_t) * n_ref, cudaMemcpyDeviceToHost, st));
        FB_CUDA(cudaStreamSynchronize(st));
    }
    return rows_common(ctx, d_ref_ids, h_ref.data(), n_ref, d_col_ids, n_col, d_out, elem_bytes, stream);
}

int famsa_lcs_rows(famsa_ctx* ctx, const uint32_t* ref_ids, uint32_t n_ref, const uint32_t* col_ids, uint32_t n_col,
                   void* out, int elem_bytes)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->lcs.n == 0) { set_error("famsa_lcs_upload has not been called"); return FAMSA_E_STATE; }
    int rc = check_elem(ctx, elem_bytes);
    if (rc) return rc;
    if (!col_ids && n_col > ctx->lcs.n) { set_error("n_col exceeds the number of sequences"); return FAMSA_E_INVALID; }
    if ((n_ref && !ref_ids) || (n_ref && n_col && !out)) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    if (col_ids)
        for (uint32_t k = 0; k < n_col; ++k)
            if (col_ids[k] >= ctx->lcs.n) { set_error("col id out of range"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(ctx->device));
    fb::LcsState& S = ctx->lcs;
    const uint64_t cells = (uint64_t)n_ref * n_col;
    rc = S.d_out_stage.reserve(std::max<uint64_t>(cells, 1) * elem_bytes);
    if (rc) return rc;
    rc = S.d_ids_a.reserve(sizeof(uint32_t) * std::max(1u, n_ref));
    if (rc) return rc;
    const uint32_t* d_cols = nullptr;
    if (col_ids) {
        rc = S.d_ids_b.reserve(sizeof(uint32_t) * std::max(1u, n_col));
        if (rc) return rc;
        FB_CUDA(cudaMemcpyAsync(S.d_ids_b.p, col_ids, sizeof(uint32_t) * n_col, cudaMemcpyHostToDevice, ctx->stream));
        d_cols = S.d_ids_b.as<uint32_t>();
    }
    if (n_ref)
        FB_CUDA(cudaMemcpyAsync(S.d_ids_a.p, ref_ids, sizeof(uint32_t) * n_ref, cudaMemcpyHostToDevice, ctx->stream));
    rc = rows_common(ctx, S.d_ids_a.as<uint32_t>(), ref_ids, n_ref, d_cols, n_col, S.d_out_stage.p, elem_bytes, nullptr);
    if (rc) return rc;
    if (cells) FB_CUDA(cudaMemcpy(out, S.d_out_stage.p, cells * elem_bytes, cudaMemcpyDeviceToHost));
    return FAMSA_OK;
}

int famsa_lcs_prim(famsa_ctx* ctx, int distance_kind, int32_t* edge_from, int32_t* edge_to, double* edge_dist,
                   int32_t* prim_order)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->lcs.n == 0) { set_error("famsa_lcs_upload has not been called"); return FAMSA_E_STATE; }
    if (!prim_order || (ctx->lcs.n > 1 && (!edge_from || !edge_to || !edge_dist))) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    if (distance_kind != 0 && distance_kind != 1) { set_error("MSTPrim is instantiated for distance_kind 0 and 1 only"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = fb::scratch_acquire(ctx, ctx->stream);
    if (rc) return rc;
    rc = fb::lcs_prim(ctx, distance_kind, edge_from, edge_to, edge_dist, prim_order);
    if (rc) return rc;
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    fb::scratch_release(ctx, ctx->stream, true);
    return finish_timing(ctx);
}

int famsa_lcs_assign(famsa_ctx* ctx, const uint32_t* seed_ids, uint32_t n_seeds, int distance_kind, uint32_t* assignments,
                     float* min_dist)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->lcs.n == 0) { set_error("famsa_lcs_upload has not been called"); return FAMSA_E_STATE; }
    if (!seed_ids || !n_seeds || !assignments || !min_dist) { set_error("NULL argument / no seeds"); return FAMSA_E_INVALID; }
    if (distance_kind < 0 || distance_kind > 2) { set_error("distance_kind must be 0, 1 or 2"); return FAMSA_E_INVALID; }
    for (uint32_t k = 0; k < n_seeds; ++k)
        if (seed_ids[k] >= ctx->lcs.n) { set_error("seed id out of range"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = fb::scratch_acquire(ctx, ctx->stream);
    if (rc) return rc;
    rc = fb::lcs_assign(ctx, seed_ids, n_seeds, distance_kind, assignments, min_dist);
    if (rc) return rc;
    FB_CUDA(cudaStreamSynchronize(ctx->stream));
    fb::scratch_release(ctx, ctx->stream, true);
    return finish_timing(ctx);
}

int famsa_lcs_upgma(famsa_ctx* ctx, int distance_kind, int modified, int32_t* tree)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->lcs.n < 2) { set_error("famsa_lcs_upgma needs at least two uploaded sequences"); return FAMSA_E_STATE; }
    if (!tree) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    if (distance_kind < 0 || distance_kind > 2) { set_error("distance_kind must be 0, 1 or 2"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = fb::scratch_acquire(ctx, ctx->stream);
    if (rc) return rc;
    rc = fb::lcs_upgma(ctx, distance_kind, modified, tree);
    if (rc) return rc;
    fb::scratch_release(ctx, ctx->stream, true);
    return finish_timing(ctx);
}

int famsa_lcs_upgma_from_triangle(famsa_ctx* ctx, int distance_kind, int modified, const void* d_triangle, int elem_bytes, int32_t* tree)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->lcs.n < 2) { set_error("famsa_lcs_upgma needs at least two uploaded sequences"); return FAMSA_E_STATE; }
    if (!tree || !d_triangle) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    if (distance_kind < 0 || distance_kind > 2) { set_error("distance_kind must be 0, 1 or 2"); return FAMSA_E_INVALID; }
    if (elem_bytes != 2 && elem_bytes != 4) { set_error("elem_bytes must be 2 or 4"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = fb::scratch_acquire(ctx, ctx->stream);
    if (rc) return rc;
    rc = fb::lcs_upgma(ctx, distance_kind, modified, tree, d_triangle, elem_bytes);
    if (rc) return rc;
    fb::scratch_release(ctx, ctx->stream, true);
    return finish_timing(ctx);
}

int famsa_lcs_assign_shard(famsa_ctx* ctx, const uint32_t* seed_ids, uint32_t n_seeds, int distance_kind, uint32_t shard,
                           uint32_t n_shards, int64_t* d_packed, void* stream)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->lcs.n == 0) { set_error("famsa_lcs_upload has not been called"); return FAMSA_E_STATE; }
    if (!seed_ids || !n_seeds || !d_packed) { set_error("NULL argument / no seeds"); return FAMSA_E_INVALID; }
    if (distance_kind < 0 || distance_kind > 2) { set_error("distance_kind must be 0, 1 or 2"); return FAMSA_E_INVALID; }
    if (!n_shards || shard >= n_shards) { set_error("shard out of range"); return FAMSA_E_INVALID; }
    for (uint32_t k = 0; k < n_seeds; ++k)
        if (seed_ids[k] >= ctx->lcs.n) { set_error("seed id out of range"); return FAMSA_E_INVALID; }
    FB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    int rc = fb::scratch_acquire(ctx, st);
    if (rc) return rc;
    rc = fb::lcs_assign_shard(ctx, seed_ids, n_seeds, distance_kind, shard, n_shards, reinterpret_cast<long long*>(d_packed), st);
    if (rc) return rc;
    if (!stream) { FB_CUDA(cudaStreamSynchronize(st)); fb::scratch_release(ctx, st, true); return finish_timing(ctx); }
    // the seed ids were handed to an async copy from pageable memory: staged by the driver before the call returned
    return fb::scratch_release(ctx, st, false);
}

int famsa_lcs_last_timing(const famsa_ctx* ctx, float* total_ms, float* main_kernel_ms, uint64_t* n_pairs)
{
    FB_CHECK_CTX(ctx);
    if (total_ms) *total_ms = ctx->lcs.last_total_ms;
    if (main_kernel_ms) *main_kernel_ms = ctx->lcs.last_main_ms;
    if (n_pairs) *n_pairs = ctx->lcs.last_pairs;
    return FAMSA_OK;
}

// ------------------------------------------------------------------ HP-2

static int dp_finish_timing(famsa_ctx* ctx)
{
    float total = 0.f, k = 0.f;
    FB_CUDA(cudaEventSynchronize(ctx->ev[3]));
    FB_CUDA(cudaEventElapsedTime(&total, ctx->ev[0], ctx->ev[3]));
    FB_CUDA(cudaEventElapsedTime(&k, ctx->ev[1], ctx->ev[2]));
    ctx->dp.last_total_ms = total;
    ctx->dp.last_kernel_ms = k;
    return FAMSA_OK;
}

int famsa_dp_align_batch(famsa_ctx* ctx, const famsa_dp_job* jobs, uint32_t n_jobs, const int64_t gaps[4],
                         famsa_dp_result* results, uint8_t* path_buf, uint8_t* dirs_buf)
{
    FB_CHECK_CTX(ctx);
    if (n_jobs && (!jobs || !gaps || !results || !path_buf)) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = fb::scratch_acquire(ctx, ctx->stream);
    if (rc) return rc;
    rc = fb::dp_run_host(ctx, jobs, n_jobs, gaps, results, path_buf, dirs_buf);
    if (rc) return rc;
    fb::scratch_release(ctx, ctx->stream, true);
    return dp_finish_timing(ctx);
}

int famsa_dp_align_batch_device(famsa_ctx* ctx, const famsa_dp_job* jobs, uint32_t n_jobs, const int64_t gaps[4],
                                famsa_dp_result* d_results, uint8_t* d_path_buf, uint8_t* d_dirs_buf, void* stream)
{
    FB_CHECK_CTX(ctx);
    if (n_jobs && (!jobs || !gaps || !d_results || !d_path_buf)) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    FB_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    int rc = fb::scratch_acquire(ctx, st);
    if (rc) return rc;
    rc = fb::dp_run_device(ctx, jobs, nullptr, n_jobs, gaps, d_results, d_path_buf, d_dirs_buf, nullptr, nullptr, st);
    if (rc) return rc;
    if (!stream) { FB_CUDA(cudaStreamSynchronize(st)); fb::scratch_release(ctx, st, true); return dp_finish_timing(ctx); }
    return fb::scratch_release(ctx, st, false);
}

int famsa_dp_last_timing(const famsa_ctx* ctx, float* total_ms, float* kernel_ms, uint64_t* n_cells)
{
    FB_CHECK_CTX(ctx);
    if (total_ms) *total_ms = ctx->dp.last_total_ms;
    if (kernel_ms) *kernel_ms = ctx->dp.last_kernel_ms;
    if (n_cells) *n_cells = ctx->dp.last_cells;
    return FAMSA_OK;
}

// ------------------------------------------------------------------ resident profiles

int famsa_prof_set_scoring(famsa_ctx* ctx, const int64_t score_matrix[24 * 24])
{
    FB_CHECK_CTX(ctx);
    if (!score_matrix) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    FB_CUDA(cudaSetDevice(ctx->device));
    return fb::prof_set_scoring(ctx, score_matrix);
}

int famsa_prof_put(famsa_ctx* ctx, const famsa_dp_profile* profiles, uint32_t n, uint32_t* ids_out)
{
    FB_CHECK_CTX(ctx);
    if (n && (!profiles || !ids_out)) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    FB_CUDA(cudaSetDevice(ctx->device));
    return fb::prof_put(ctx, profiles, n, ids_out);
}

int famsa_prof_merge_batch(famsa_ctx* ctx, const famsa_prof_merge* merges, uint32_t n, const int64_t gaps[4],
                           uint32_t* merged_ids_out, famsa_dp_result* results, uint8_t* path_buf, uint64_t path_cap)
{
    FB_CHECK_CTX(ctx);
    if (n && (!merges || !gaps || !merged_ids_out || !results || !path_buf)) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = fb::scratch_acquire(ctx, ctx->stream);
    if (rc) return rc;
    rc = fb::prof_merge_batch(ctx, merges, n, gaps, merged_ids_out, results, path_buf, path_cap);
    if (rc || !n) return rc;
    return dp_finish_timing(ctx);
}

int famsa_prof_align_tree(famsa_ctx* ctx, const int32_t* tree, uint32_t n_leaves, const int64_t gaps[4],
                          famsa_dp_result* results, uint32_t* root_id_out, uint64_t* path_bytes_out, famsa_tree_stats* stats)
{
    FB_CHECK_CTX(ctx);
    if (!tree || !gaps || !results) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    FB_CUDA(cudaSetDevice(ctx->device));
    int rc = fb::scratch_acquire(ctx, ctx->stream);
    if (rc) return rc;
    rc = fb::prof_align_tree(ctx, tree, n_leaves, gaps, results, root_id_out, path_bytes_out, stats);
    if (rc) return rc;
    fb::scratch_release(ctx, ctx->stream, true);
    return FAMSA_OK;
}

int famsa_prof_tree_paths(famsa_ctx* ctx, uint8_t* path_buf, uint64_t path_cap)
{
    FB_CHECK_CTX(ctx);
    if (!path_buf) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    return fb::prof_tree_paths(ctx, path_buf, path_cap);
}

int famsa_prof_get(famsa_ctx* ctx, uint32_t id, uint32_t* width, uint32_t* card, int64_t* scores, int32_t* counters)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    FB_CUDA(cudaSetDevice(ctx->device));
    return fb::prof_get(ctx, id, width, card, scores, counters);
}

int famsa_prof_drop(famsa_ctx* ctx, const uint32_t* ids, uint32_t n)
{
    FB_CHECK_CTX(ctx);
    if (n && !ids) { set_error("NULL argument"); return FAMSA_E_INVALID; }
    std::lock_guard<std::mutex> lk(ctx->mu);
    FB_CUDA(cudaSetDevice(ctx->device));
    return fb::prof_drop(ctx, ids, n);
}

int famsa_prof_last_timing(famsa_ctx* ctx, float* total_ms, float* construct_ms)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    FB_CUDA(cudaSetDevice(ctx->device));
    return fb::prof_last_timing(ctx, total_ms, construct_ms);
}

int famsa_prof_stats(famsa_ctx* ctx, uint64_t* n_live, uint64_t* resident_bytes)
{
    FB_CHECK_CTX(ctx);
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (n_live) *n_live = ctx->prof.n_live;
    if (resident_bytes) *resident_bytes = ctx->prof.resident_bytes;
    return FAMSA_OK;
}

// ------------------------------------------------------------------ Transform<T, Distance>
// reference src/tree/AbstractTreeGenerator.hpp:28-82; kept on the host so that distances are
// bit-identical to the reference's (same libm pow, same float/double narrowing points).

double famsa_transform_f64(int kind, uint32_t lcs, uint32_t len1, uint32_t len2)
{
    if (kind == 2) return (double)lcs / std::min(len1, len2);
    const double indel = (double)(len1 + len2 - 2 * lcs);
    if (!lcs) return std::nextafter(DBL_MAX, 0.0);
    if (kind == 0) return (double)std::pow((double)(uint32_t)indel, 0.75) / (double)lcs;
    return indel / lcs;
}

float famsa_transform_f32(int kind, uint32_t lcs, uint32_t len1, uint32_t len2)
{
    if (kind == 2) return (float)lcs / std::min(len1, len2);
    const float indel = (float)(len1 + len2 - 2 * lcs);
    if (!lcs) return (float)std::nextafter((double)FLT_MAX, 0.0);
    if (kind == 0) return (float)std::pow((double)(uint32_t)indel, 0.75) / (float)lcs;
    return indel / lcs;
}

} // extern "C"
