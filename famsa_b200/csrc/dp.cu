// HP-2: profile-profile / sequence-profile / sequence-sequence affine-gap DP with position-specific
// gap costs, many independent merges per launch, plus on-device traceback.
//
// Replaces CProfile::Align and its cell loops (reference src/core/profile.cpp:244-305,
// profile_seq.cpp:24-892, profile_par.cpp:26-903), DP_SolveGapsProblemWhenStarting/Continuing
// (profile.cpp:1223-1315) and the traceback of ConstructProfile (profile.cpp:727-782).
// Semantics follow SURVEY.md Appendix B; arithmetic is int64 with NEG = -(1<<62) unsaturated.
//
// Design (three kernels per batch):
//   k_dp_prep  one block per merge: variant + orientation (CProfile::Align), the column-side constant
//              records (gap scores + gap-correction counts, 8 x int64 per column, structure of arrays), row 0 of
//              the DP (block-parallel prefix sum) and, for the Seq* variants, each row's residue plus a transposed
//              copy of the column profile's scores.
//   k_dp_t     embarrassingly parallel: T[i][j] = sum_k counters_row[i][k] * scores_col[j][k], the column-
//              pair score that does not depend on the DP state (profile_par.cpp:695-711).  Taking this dot
//              product out of the wavefront removes it from the latency-critical loop.  One block = one 32-row
//              stripe x 128 columns, column scores in registers, row counters broadcast; the result is stored
//              SKEWED (stripe, wavefront step, lane) so that the fill kernel reads it with one 256-byte load per step.
//   k_dp_fill  the recurrence itself: a team of warps per merge, 32-row stripes, lane L owns row i0+L and
//              at step s computes column s-L (anti-diagonal wavefront); (D,H,V) of the cell above arrives
//              by warp shuffle, the left neighbour stays in registers, T of the next step is prefetched into a
//              register, the column records live in a per-warp shared-memory ring filled by cp.async one chunk
//              ahead, and the direction bytes go out skewed as well (one 32-byte store per step).  Stripes of one
//              merge run as a lock-step staircase over the warps of the block (or of a thread-block cluster for
//              very wide merges), handing the boundary row over through L2.  The same kernel then walks the
//              direction matrix back and emits the path.
//   k_dp_unskew only when the caller asks for CDPMatrix bytes: skewed directions -> row-major.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <thread>

#include <cooperative_groups.h>

#include "ctx.h"

namespace fb {

constexpr int kGO = 25, kGE = 26, kTE = 27, kTO = 28;   // GAP_OPEN, GAP_EXT, GAP_TERM_EXT, GAP_TERM_OPEN (defs.h:62-66)
constexpr long long kNeg = -(1ll << 62);
constexpr int kDpWarps = 4;               // block size of the one-warp-per-merge fill kernel
constexpr int kDpTeamWarps = 8;           // warps cooperating on one large merge
constexpr uint32_t kDpTeamMinWidth = 96;  // min(w1, w2) above which a merge gets a team
constexpr int kDpCluster = 8;             // thread blocks per cluster for the widest merges
constexpr uint32_t kDpClusterMinWidth = 1024;   // min(w1, w2) above which a merge gets a whole cluster
constexpr int kTThreads = 256, kTCellsPerThread = 4;

struct DpJobDev {
    const long long* s1; const int* c1;
    const long long* s2; const int* c2;
    uint32_t w1, card1, w2, card2;
    unsigned long long path_off, dirs_off, scratch_off, t_off;   // t_off: element offset of the job's skewed T / directions
};

struct DpMeta {            // written by k_dp_prep
    const long long* SR; const int* CR;
    const long long* SC; const int* CC;
    uint32_t WR, WC;
    int nR, nC, var, sw;
    int bad;               // a count of the ProfProf tables is negative or exceeds the member count: not a profile CProfile can build
    int narrow;            // every score of the column profile fits in int32: k_dp_t multiplies 32 x 32 -> 64
};

// per column j of the column profile, 64 bytes, meaning depends on the variant:
//   ProfProf: gap scores {S[j][GO], S[j][GE], S[j][TO], S[j][TE]}, chg, b0 = (s_o, s_e), b1 = (s_to, s_te), b2 = (k_e, k_te)
//   SeqProf : same gap scores, chg, b0 = gcv1, b1 = contv1 (profile_par.cpp:204-211)
//   SeqSeq  : unused
struct ColInfo {
    long long cgo, cge, cto, cte;
    long long chg;
    long long b0, b1, b2;
};
static_assert(sizeof(ColInfo) == 64, "ColInfo layout");

struct RowNz {             // per row of a Seq* merge: k[0] = its residue (the list form served an earlier sparse k_dp_t), 160 bytes
    int n;
    int c[30];
    unsigned char k[30];
    unsigned char pad[6];
};
static_assert(sizeof(RowNz) == 160, "RowNz layout");

struct Cell { long long D, H, V, pad; };   // 32 bytes: two 16-byte cp.async / st.cg.v2 transfers
static_assert(sizeof(Cell) == 32, "Cell layout");
constexpr int kChunk = 8;                   // boundary-row columns handed over per cp.async batch (one macro step)

__host__ __device__ inline unsigned long long align_up(unsigned long long v, unsigned long long a) { return (v + a - 1) / a * a; }

// Skewed (wavefront-major) storage of per-cell data: stripe k (rows 32k+1 .. 32k+32), wavefront step s, lane l hold
// cell (32k+1+l, s-l).  One warp step of k_dp_fill then touches 32 consecutive elements (one 256-byte T read, one
// 32-byte direction write) instead of 32 different rows.
__host__ __device__ inline unsigned long long skew_elems_oriented(uint32_t wr, uint32_t wc)
{
    return (unsigned long long)((wr + 31) / 32) * 32ull * ((unsigned long long)wc + 32);
}
// the orientation is chosen on the device (k_dp_prep): reserve for the larger of the two
__host__ __device__ inline unsigned long long skew_elems(uint32_t w1, uint32_t w2)
{
    const unsigned long long a = skew_elems_oriented(w1, w2), b = skew_elems_oriented(w2, w1);
    return a > b ? a : b;
}

constexpr int kColFields = 8;               // ColInfo as structure-of-arrays: field f of column j at col[f * cstride + j]
constexpr int kRing = 64;                   // shared-memory window of column records per warp (columns mod 64)

// scratch layout of one job (all sections 128-byte aligned)
struct Scratch {
    unsigned long long col, brow, rownz, s2t, tmp, lastv, total, cstride;
    __host__ __device__ Scratch(uint32_t w1, uint32_t w2)
    {
        const unsigned long long wm = (w1 > w2 ? w1 : w2) + 1ull;
        cstride = align_up(wm + 1, 16);
        col = 0;
        brow = align_up(col + 8ull * kColFields * cstride, 128);
        rownz = align_up(brow + sizeof(Cell) * wm, 128);
        s2t = align_up(rownz + sizeof(RowNz) * wm, 128);
        tmp = align_up(s2t + 8ull * 30 * wm, 128);
        lastv = align_up(tmp + w1 + w2, 128);
        total = lastv + 128;
    }
};

__device__ __forceinline__ long long pack2(int lo, int hi) { return (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo); }
__device__ __forceinline__ int lo32(long long v) { return (int)(unsigned long long)v; }
__device__ __forceinline__ int hi32(long long v) { return (int)((unsigned long long)v >> 32); }
// The counts that weight the gap scores (how many sequences open / extend a gap, how many hold a residue) are
// non-negative for every profile CProfile can build; k_dp_prep verifies that (DpMeta::bad) so that the cell loop may
// multiply  int64 score x uint32 count  with two instructions (IMAD.WIDE.U32 + IMAD) instead of the signed 64 x 64 form.
__device__ __forceinline__ unsigned ulo32(long long v) { return (unsigned)(unsigned long long)v; }
__device__ __forceinline__ unsigned uhi32(long long v) { return (unsigned)((unsigned long long)v >> 32); }

__device__ __forceinline__ long long shfl_up_ll(long long v)
{
    int lo = lo32(v), hi = hi32(v);
    lo = __shfl_up_sync(0xffffffffu, lo, 1);
    hi = __shfl_up_sync(0xffffffffu, hi, 1);
    return pack2(lo, hi);
}

__device__ __forceinline__ long long shfl_up_ll_by(long long v, int delta)
{
    int lo = lo32(v), hi = hi32(v);
    lo = __shfl_up_sync(0xffffffffu, lo, delta);
    hi = __shfl_up_sync(0xffffffffu, hi, delta);
    return pack2(lo, hi);
}

// boundary-row traffic goes through L2 (.cg): written by lane 31 of one stripe, read by lane 0 of the next
__device__ __forceinline__ void store_cell(Cell* p, const Cell& c)
{
    longlong2* q = reinterpret_cast<longlong2*>(p);
    __stcg(q, make_longlong2(c.D, c.H));
    __stcg(q + 1, make_longlong2(c.V, 0));
}
// global (L2) -> shared without staging registers: the warp does not wait for the data
__device__ __forceinline__ void cp_async_cell(Cell* smem_dst, const Cell* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16), "l"(reinterpret_cast<const char*>(gsrc) + 16) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
constexpr int kPrefetch = 8;                 // columns of look-ahead for the L1 prefetches
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// a if a>b && a>c; else b if b>c; else c  (strict comparisons, fixed priority)
__device__ __forceinline__ int pick3(long long a, long long b, long long c, int da, int db, int dc, long long& out)
{
    const bool aw = (a > b) & (a > c);
    const bool bw = b > c;
    const long long bc = bw ? b : c;
    const int dbc = bw ? db : dc;
    out = aw ? a : bc;
    return aw ? da : dbc;
}
// a value that can never win a strict comparison: turns a 3-way pick into the reference's 2-way form
constexpr long long kNever = (long long)0x8000000000000000ull;

// DP_SolveGapsProblemWhenStarting / WhenContinuing (profile.cpp:1223-1315) for column c
__device__ __forceinline__ void solve_gaps(const int* __restrict__ cnt, uint32_t c, uint32_t width, int card,
                                           int& s_o, int& s_e, int& s_to, int& s_te, int& k_e, int& k_te)
{
    const int* cc = cnt + (size_t)c * 32;
    if (c >= width) {
        s_te = cc[kTO] + cc[kTE]; s_to = card - s_te; s_o = 0; s_e = 0;
        k_te = card; k_e = 0;
    } else {
        const int* cn = cc + 32;
        s_to = cn[kTO]; s_te = cc[kTO] + cc[kTE]; s_e = cc[kGO] + cc[kGE];
        s_o = card - s_e - s_to - s_te;
        k_te = cn[kTO] + cc[kTO] + cc[kTE]; k_e = card - k_te;
    }
}

__device__ __forceinline__ int seq_symbol(const int* __restrict__ cnt, uint32_t c)
{
    const int* cc = cnt + (size_t)c * 32;
    for (int k = 0; k < 24; ++k)
        if (cc[k]) return k;
    return 22;
}

struct DpParams {
    const DpJobDev* jobs;
    DpMeta* meta;
    const uint32_t* order;        // launch slot -> job index
    uint32_t n_jobs;              // jobs of this launch (a sub-batch, or one fill class of it)
    uint32_t job_base;            // first job id of the sub-batch (k_dp_prep / k_dp_t index jobs as job_base + x)
    long long go, ge, to, te;
    unsigned char* dirs;          // caller-visible row-major direction matrices (CDPMatrix layout) or nullptr
    unsigned char* sdirs;         // internal skewed direction bytes of the sub-batch
    unsigned char* path;          // all paths (forward order)
    unsigned char* scratch;
    long long* T;                 // all T matrices
    const unsigned long long* tblock;   // k_dp_unskew: first block of each job (n_jobs + 1 entries)
    const unsigned long long* t2block;  // k_dp_t: first block of each job (n_jobs + 1 entries)
    famsa_dp_result* results;
};

// ------------------------------------------------------------------------------------------------
// k_dp_prep: one block per job
// ------------------------------------------------------------------------------------------------
constexpr int kPrepThreads = 512;
__global__ void __launch_bounds__(kPrepThreads) k_dp_prep(const DpParams P)
{
    __shared__ unsigned long long sm_nz[2];
    const uint32_t jid = P.job_base + blockIdx.x;
    const DpJobDev J = P.jobs[jid];
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    if (tid < 2) sm_nz[tid] = 0;
    __syncthreads();

    // variant and orientation (CProfile::Align, profile.cpp:254-304)
    int var, sw = 0;
    if (J.card1 == 1 && J.card2 == 1) var = 0;
    else if (J.card1 == 1) var = 1;
    else if (J.card2 == 1) { var = 1; sw = 1; }
    else {
        var = 2;
        unsigned long long nz1 = 0, nz2 = 0;
        for (size_t k = tid; k < ((size_t)J.w1 + 1) * 32; k += nthr) nz1 += J.c1[k] != 0;
        for (size_t k = tid; k < ((size_t)J.w2 + 1) * 32; k += nthr) nz2 += J.c2[k] != 0;
        for (int o = 16; o; o >>= 1) {
            nz1 += __shfl_xor_sync(0xffffffffu, nz1, o);
            nz2 += __shfl_xor_sync(0xffffffffu, nz2, o);
        }
        if ((tid & 31) == 0) { atomicAdd(&sm_nz[0], nz1); atomicAdd(&sm_nz[1], nz2); }
        __syncthreads();
        if (!(sm_nz[0] * (unsigned long long)J.w2 < sm_nz[1] * (unsigned long long)J.w1)) sw = 1;
    }
    const long long* SR = sw ? J.s2 : J.s1;  const int* CR = sw ? J.c2 : J.c1;
    const long long* SC = sw ? J.s1 : J.s2;  const int* CC = sw ? J.c1 : J.c2;
    const uint32_t WR = sw ? J.w2 : J.w1, WC = sw ? J.w1 : J.w2;
    const int nR = (int)(sw ? J.card2 : J.card1), nC = (int)(sw ? J.card1 : J.card2);
    // ProfProf: every count the cell loop multiplies with must be >= 0 (see ulo32)
    __shared__ int sm_bad, sm_wide;
    if (tid == 0) { sm_bad = 0; sm_wide = 0; }
    __syncthreads();
    if (var == 2) {
        int bad = 0, wide = 0;
        for (size_t e = 32 + tid; e < ((size_t)WC + 1) * 32; e += nthr) {       // columns 1..WC, rows 0..29 feed k_dp_t
            const long long v = SC[e];
            wide |= (e & 31) < 30 && v != (long long)(int)v;
        }
        if (wide) sm_wide = 1;
        for (int side = 0; side < 2; ++side) {
            const int* cnt = side ? CC : CR;
            const uint32_t w = side ? WC : WR;
            const int card = side ? nC : nR;
            for (uint32_t c = 1 + tid; c <= w; c += nthr) {
                int a, b, d, e, f, g;
                solve_gaps(cnt, c, w, card, a, b, d, e, f, g);
                bad |= (a | b | d | e | f | g) < 0;
                const int* cc = cnt + (size_t)c * 32;
                int neg = 0, over = 0;
                for (int k = 0; k < 30; ++k) { neg |= cc[k]; over |= cc[k] > card; }
                bad |= neg < 0 || over;
            }
        }
        if (bad) sm_bad = 1;
    }
    __syncthreads();
    if (tid == 0) {
        DpMeta m;
        m.SR = SR; m.CR = CR; m.SC = SC; m.CC = CC; m.WR = WR; m.WC = WC; m.nR = nR; m.nC = nC; m.var = var; m.sw = sw;
        m.bad = sm_bad;
        m.narrow = var == 2 && !sm_wide;
        P.meta[jid] = m;
    }
    const Scratch L(J.w1, J.w2);
    unsigned char* scratch = P.scratch + J.scratch_off;
    long long* col = reinterpret_cast<long long*>(scratch + L.col);
    Cell* brow = reinterpret_cast<Cell*>(scratch + L.brow);
    RowNz* rownz = reinterpret_cast<RowNz*>(scratch + L.rownz);
    long long* s2t = reinterpret_cast<long long*>(scratch + L.s2t);
    const long long go = P.go, ge = P.ge, to = P.to, te = P.te;
    const size_t ldc = (size_t)WC + 1;

    // column records (structure of arrays)
    for (uint32_t j = tid; j <= WC; j += nthr) {
        ColInfo ci = {0, 0, 0, 0, 0, 0, 0, 0};
        if (j >= 1 && var != 0) {
            int s_o, s_e, s_to, s_te, k_e, k_te;
            solve_gaps(CC, j, WC, nC, s_o, s_e, s_to, s_te, k_e, k_te);
            const int* cc = CC + (size_t)j * 32;
            const long long* sc = SC + (size_t)j * 32;
            ci.cgo = sc[kGO]; ci.cge = sc[kGE]; ci.cto = sc[kTO]; ci.cte = sc[kTE];
            ci.chg = (long long)cc[kGO] * (ge - go) + (long long)cc[kTO] * (te - to);
            if (var == 2) { ci.b0 = pack2(s_o, s_e); ci.b1 = pack2(s_to, s_te); ci.b2 = pack2(k_e, k_te); }
            else { ci.b0 = go * s_o + ge * s_e + to * s_to + te * s_te; ci.b1 = ge * k_e + te * k_te; }
        }
        const long long v[kColFields] = {ci.cgo, ci.cge, ci.cto, ci.cte, ci.chg, ci.b0, ci.b1, ci.b2};
#pragma unroll
        for (int f = 0; f < kColFields; ++f) col[(size_t)f * L.cstride + j] = v[f];
    }
    // row 0 (profile_par.cpp:531-555; SeqSeq profile_seq.cpp:48-69): H(0, j) is a running sum over the columns --
    // every thread sums a contiguous segment, the segment totals are combined through shared memory
    {
        __shared__ long long sm_seg[kPrepThreads];
        auto term = [&](uint32_t j) -> long long {
            const long long* sc = SC + (size_t)j * 32;
            if (var == 0) return j == 1 ? to : te;             // max(H, D = NEG) + te
            if (var == 1) return j == 1 ? sc[kTO] : sc[kTE];
            return (j == 1 ? sc[kTO] : sc[kTE]) * nR;
        };
        const uint32_t seg = (WC + nthr - 1) / nthr;
        const uint32_t ja = 1 + tid * seg, jb = ja + seg - 1 < WC ? ja + seg - 1 : WC;
        long long sum = 0;
        for (uint32_t j = ja; j <= jb; ++j) sum += term(j);
        sm_seg[tid] = sum;
        __syncthreads();
        long long h = 0;
        for (uint32_t u = 0; u < tid; ++u) h += sm_seg[u];
        for (uint32_t j = ja; j <= jb; ++j) {
            h += term(j);
            store_cell(brow + j, Cell{kNeg, j == WC ? kNeg : h, kNeg, 0});
        }
        if (tid == 0) store_cell(brow, Cell{0, kNeg, kNeg, 0});
    }
    if (var == 2) return;       // k_dp_t reads the ProfProf tables directly
    // Seq* variants: the residue of every row, and a transposed copy of the column profile's scores (s2t[k][j], k < 30)
    // so that k_dp_t's single look-up per cell is coalesced
    for (uint32_t i = tid; i <= WR; i += nthr) {
        RowNz* dst = rownz + i;
        dst->n = i >= 1;
        if (i >= 1) { dst->c[0] = 1; dst->k[0] = (unsigned char)seq_symbol(CR, i); }
    }
    for (size_t e = tid; e < ldc * 32; e += nthr) {
        const uint32_t j = (uint32_t)(e / 32), k = (uint32_t)(e % 32);
        if (k < 30) s2t[(size_t)k * ldc + j] = SC[e];
    }
}

// ------------------------------------------------------------------------------------------------
// k_dp_t: T[i][j] = sum_k counters_row[i][k] * scores_col[j][k] for every cell of every job, written in the
// skewed layout k_dp_fill streams.  A block owns one 32-row stripe x 128 columns of one merge: every thread keeps
// its column's 30 scores in registers (as 32-bit halves) and walks down the 32 rows; the row's counters are the
// same for the whole block (broadcast 128-bit loads), so a cell costs 30 x (IMAD.WIDE + IMAD) and a fraction of a
// load wavefront.  The tile goes through shared memory so that the skewed store is made of full 256-byte rows.
// Seq* variants (one residue per row) read the single score they need from the transposed copy instead.
// ------------------------------------------------------------------------------------------------
constexpr int kTCols = 128;                 // columns (= threads) per k_dp_t block

__host__ __device__ inline unsigned long long t_blocks_oriented(uint32_t wr, uint32_t wc)
{
    return (unsigned long long)((wc + kTCols - 1) / kTCols) * ((wr + 31) / 32);
}
__host__ __device__ inline unsigned long long t_blocks(uint32_t w1, uint32_t w2)
{
    const unsigned long long a = t_blocks_oriented(w1, w2), b = t_blocks_oriented(w2, w1);
    return a > b ? a : b;
}

// The usual ProfProf case -- every score fits in int32 and the row profile has at most 127 (NDA = 1) or 32767
// (NDA = 2) members, so every counter is one or two byte digits: T = C x S^T is an exact integer GEMM with K = 32
// symbols, done on the tensor cores.  Scores are split into four byte digits (three unsigned, the top one signed);
// IMMA.16832 accumulates each digit plane in int32 (30 x 255 x 255 < 2^21) and the planes are recombined with shifts in
// 64 bits.  Warp w owns columns j0 + 32w .. +31 (four 8-column tiles) and both 16-row tiles of the stripe.
template <int NDA>
__device__ __forceinline__ void t_tile_mma(const DpMeta& M, uint32_t i0, uint32_t nrows, uint32_t j0, uint32_t WC,
                                           long long (*tile)[kTCols + 2])
{
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
    unsigned afrag[NDA][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 4; ++h) {                              // h&1: row +8, h>>1: symbols 16..31
            const uint32_t row = (uint32_t)mt * 16 + g + (h & 1) * 8, k = (h >> 1) * 16 + t4 * 4;
            unsigned lo = 0, hi = 0;
            if (row < nrows) {
                const int4 c = *reinterpret_cast<const int4*>(M.CR + (size_t)(i0 + row) * 32 + k);
                const unsigned x = (unsigned)c.x, y = (unsigned)c.y, z = (unsigned)c.z, w = (unsigned)c.w;
                lo = (x & 0xffu) | (y & 0xffu) << 8 | (z & 0xffu) << 16 | (w & 0xffu) << 24;
                hi = (x >> 8 & 0xffu) | (y >> 8 & 0xffu) << 8 | (z >> 8 & 0xffu) << 16 | (w >> 8 & 0xffu) << 24;
                if (k == 28) { lo &= 0xffffu; hi &= 0xffffu; }     // rows 30 (GAP) and 31 (GUARD) are not part of the sum
            }
            afrag[0][mt][h] = lo;
            if (NDA == 2) afrag[NDA - 1][mt][h] = hi;
        }
#pragma unroll 1
    for (int nt = 0; nt < 4; ++nt) {
        const uint32_t cb = warp * 32 + (uint32_t)nt * 8;           // tile-local column of this 8-column tile
        const uint32_t jc = j0 + cb + g;                            // the column whose scores this lane supplies
        const long long* sc = M.SC + (size_t)(jc <= WC ? jc : 1) * 32;
        unsigned bfrag[4][2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const longlong2 p = *reinterpret_cast<const longlong2*>(sc + h * 16 + t4 * 4);
            const longlong2 q = *reinterpret_cast<const longlong2*>(sc + h * 16 + t4 * 4 + 2);
            const unsigned v0 = (unsigned)p.x, v1 = (unsigned)p.y, v2 = (unsigned)q.x, v3 = (unsigned)q.y;
            const unsigned t01 = __byte_perm(v0, v1, 0x5140), t23 = __byte_perm(v2, v3, 0x5140);   // bytes 0,1 interleaved
            const unsigned u01 = __byte_perm(v0, v1, 0x7362), u23 = __byte_perm(v2, v3, 0x7362);   // bytes 2,3 interleaved
            bfrag[0][h] = __byte_perm(t01, t23, 0x5410); bfrag[1][h] = __byte_perm(t01, t23, 0x7632);
            bfrag[2][h] = __byte_perm(u01, u23, 0x5410); bfrag[3][h] = __byte_perm(u01, u23, 0x7632);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            long long out[4] = {0, 0, 0, 0};
#pragma unroll
            for (int da = 0; da < NDA; ++da) {
                int acc[4][4];
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    acc[d][0] = acc[d][1] = acc[d][2] = acc[d][3] = 0;
                    if (d < 3)
                        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                     : "+r"(acc[d][0]), "+r"(acc[d][1]), "+r"(acc[d][2]), "+r"(acc[d][3])
                                     : "r"(afrag[da][mt][0]), "r"(afrag[da][mt][1]), "r"(afrag[da][mt][2]), "r"(afrag[da][mt][3]),
                                       "r"(bfrag[d][0]), "r"(bfrag[d][1]));
                    else
                        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                     : "+r"(acc[d][0]), "+r"(acc[d][1]), "+r"(acc[d][2]), "+r"(acc[d][3])
                                     : "r"(afrag[da][mt][0]), "r"(afrag[da][mt][1]), "r"(afrag[da][mt][2]), "r"(afrag[da][mt][3]),
                                       "r"(bfrag[d][0]), "r"(bfrag[d][1]));
                }
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    out[c] += ((long long)acc[0][c] + ((long long)acc[1][c] << 8) + ((long long)acc[2][c] << 16) + ((long long)acc[3][c] << 24)) << (8 * da);
            }
            // c0,c1: row g, columns 2*t4, 2*t4+1;  c2,c3: row g+8
            *reinterpret_cast<longlong2*>(&tile[mt * 16 + g][cb + t4 * 2]) = make_longlong2(out[0], out[1]);
            *reinterpret_cast<longlong2*>(&tile[mt * 16 + g + 8][cb + t4 * 2]) = make_longlong2(out[2], out[3]);
        }
    }
}

__global__ void __launch_bounds__(kTCols) k_dp_t(const DpParams P)
{
    __shared__ __align__(16) long long tile[32][kTCols + 2];     // +2: rows 16 bytes apart in the banks
    // which job does this block belong to?
    uint32_t lo = 0, hi = P.n_jobs;
    const unsigned long long b = blockIdx.x;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) / 2;
        if (P.t2block[mid] <= b) lo = mid; else hi = mid;
    }
    const uint32_t jid = P.job_base + lo;
    const DpJobDev J = P.jobs[jid];
    const DpMeta M = P.meta[jid];
    const uint32_t WR = M.WR, WC = M.WC;
    const uint32_t ctiles = (WC + kTCols - 1) / kTCols;
    const uint32_t bj = (uint32_t)(b - P.t2block[lo]);
    if (bj >= ctiles * ((WR + 31) / 32)) return;                    // reserved for the other orientation
    const uint32_t stripe = bj / ctiles;
    const uint32_t j0 = 1 + (bj % ctiles) * kTCols;                 // first column of the tile
    const uint32_t j = j0 + threadIdx.x;
    const uint32_t i0 = 1 + stripe * 32;
    const uint32_t nrows = WR - i0 + 1 < 32 ? WR - i0 + 1 : 32;
    const bool ok = j <= WC;
    if (M.var == 2 && M.narrow && M.nR <= 127) {
        t_tile_mma<1>(M, i0, nrows, j0, WC, tile);
    } else if (M.var == 2 && M.narrow && M.nR <= 32767) {
        t_tile_mma<2>(M, i0, nrows, j0, WC, tile);
    } else if (M.var == 2) {
        unsigned slo[30], shi[30];
        {
            const longlong2* sc = reinterpret_cast<const longlong2*>(M.SC + (size_t)(ok ? j : 1) * 32);
#pragma unroll
            for (int k = 0; k < 15; ++k) {
                const longlong2 v = sc[k];
                slo[2 * k] = (unsigned)v.x; shi[2 * k] = (unsigned)((unsigned long long)v.x >> 32);
                slo[2 * k + 1] = (unsigned)v.y; shi[2 * k + 1] = (unsigned)((unsigned long long)v.y >> 32);
            }
        }
        for (uint32_t l = 0; l < nrows; ++l) {
            const int4* rc = reinterpret_cast<const int4*>(M.CR + (size_t)(i0 + l) * 32);   // same address in every thread
            unsigned long long acc = 0;
            unsigned acch = 0;
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
                const int4 c = rc[k4];
                const unsigned cv[4] = {(unsigned)c.x, (unsigned)c.y, (unsigned)c.z, (unsigned)c.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = 4 * k4 + u;
                    if (k < 30) {                                   // counters are >= 0: 32 x 64 -> 64 bits, wrapping
                        acc += (unsigned long long)cv[u] * slo[k];
                        acch += cv[u] * shi[k];
                    }
                }
            }
            tile[l][threadIdx.x] = (long long)(acc + ((unsigned long long)acch << 32));
        }
    } else {
        const Scratch L(J.w1, J.w2);
        const unsigned char* scratch = P.scratch + J.scratch_off;
        const RowNz* rownz = reinterpret_cast<const RowNz*>(scratch + L.rownz);
        const long long* s2t = reinterpret_cast<const long long*>(scratch + L.s2t);
        const size_t ldc = (size_t)WC + 1;
        for (uint32_t l = 0; l < nrows; ++l) tile[l][threadIdx.x] = ok ? s2t[(size_t)rownz[i0 + l].k[0] * ldc + j] : 0;
    }
    __syncthreads();
    // skewed store: wavefront step s of this stripe holds cells (i0 + l, s - l); the tile covers steps j0 .. j0+127+31
    long long* Tk = P.T + J.t_off + (size_t)stripe * 32 * ((size_t)WC + 32);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t ds = warp; ds < kTCols + 31; ds += kTCols / 32) {
        const int c = (int)ds - (int)lane;                           // column inside the tile
        if (c >= 0 && c < kTCols && lane < nrows && j0 + (uint32_t)c <= WC)
            Tk[(size_t)(j0 + ds) * 32 + lane] = tile[lane][c];
    }
}

// Caller-visible CDPMatrix bytes (row-major, row 0 all-H, column 0 all-V) from the skewed internal directions.
__device__ __forceinline__ unsigned char dir_at(const unsigned char* __restrict__ sdirs, size_t steps, uint32_t i, uint32_t j)
{
    if (i == 0) return j ? (unsigned char)(1 | 1 << 2 | 1 << 4) : 0;
    const uint32_t q = i - 1, l = q & 31;
    return __ldcg(sdirs + ((size_t)(q >> 5) * steps + (j + l)) * 32 + l);
}

__global__ void __launch_bounds__(kTThreads) k_dp_unskew(const DpParams P)
{
    uint32_t lo = 0, hi = P.n_jobs;
    const unsigned long long b = blockIdx.x;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) / 2;
        if (P.tblock[mid] <= b) lo = mid; else hi = mid;
    }
    const uint32_t jid = P.job_base + lo;
    const DpJobDev J = P.jobs[jid];
    const DpMeta M = P.meta[jid];
    const size_t ldc = (size_t)M.WC + 1, steps = (size_t)M.WC + 32;
    const size_t cells = ((size_t)M.WR + 1) * ldc;
    const unsigned char* sd = P.sdirs + J.t_off;
    unsigned char* out = P.dirs + J.dirs_off;
    const size_t base = (size_t)(b - P.tblock[lo]) * kTThreads * kTCellsPerThread;
#pragma unroll
    for (int u = 0; u < kTCellsPerThread; ++u) {
        const size_t c = base + (size_t)u * kTThreads + threadIdx.x;
        if (c >= cells) break;
        const uint32_t i = (uint32_t)c / (uint32_t)ldc, j = (uint32_t)c - i * (uint32_t)ldc;
        out[c] = dir_at(sd, steps, i, j);
    }
}

// ------------------------------------------------------------------------------------------------
// k_dp_fill: the recurrence + traceback
// ------------------------------------------------------------------------------------------------

// One team = NW warps (x CL thread blocks of a cluster) working on one merge.  Stripe k (rows 32k+1 .. 32k+32)
// belongs to team warp k % (NW*CL).  The
// stripes of a team run as a lock-step staircase over MACRO STEPS of kChunk wavefront steps: stripe k starts
// kLag macro steps after stripe k-1, which is exactly late enough for every boundary-row column it is about
// to read (and the chunk it prefetches for the next macro step) to have been parked by lane 31 of stripe k-1.
// Because the schedule is a closed form, nobody polls: one __syncthreads per macro step orders the hand-over.
// Macro steps between consecutive stripes for a chunk of CH columns.  The consumer requests chunk `off` right after
// the barrier that opens its macro step off-1; by then the producer (kLag macro steps ahead) must have parked column
// off*CH + CH-1, which its lane 31 computes at wavefront step off*CH + CH-1 + 31:  kLag >= 3 + floor(30 / CH).
// CH = 8 (48 columns of lag) for every team size: the time of a wide merge is stripes x lag x step time, and the extra
// barriers cost less than the 16 columns of lag that CH = 16 would add (measured on the bench batch and on whole trees).
__host__ __device__ constexpr int lag_of(int ch) { return 3 + 30 / ch; }

template <int VAR, int NW, int CL, int CH>
__device__ __forceinline__ void dp_stripes(const DpParams& P, const DpMeta& M, const long long* __restrict__ T,
                                           const long long* __restrict__ col, uint32_t cstride, Cell* __restrict__ brow,
                                           unsigned char* __restrict__ dirs, uint32_t team_warp,
                                           long long* last_out, Cell (*sb)[CH], long long (*ring)[kRing])
{
    constexpr int kChunk = CH, kLag = lag_of(CH);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t WR = M.WR, WC = M.WC;
    const long long go = P.go, ge = P.ge, to = P.to, te = P.te;
    const uint32_t n_stripes = (WR + 31) / 32;
    const uint32_t steps = WC + 1 + 31;                              // wavefront steps per stripe
    const uint32_t S = (steps + kChunk - 1) / kChunk;                // macro steps per stripe
    constexpr uint32_t TW = (uint32_t)NW * CL;                       // warps in the team (CL > 1: a thread-block cluster)
    const uint32_t period = TW == 1 ? S : (S > TW * kLag ? S : TW * kLag);
    const uint32_t rounds = (n_stripes + TW - 1) / TW;
    // stripe k = r*NW + w starts at macro step r*period + w*kLag; the last stripe ends at m_end
    const uint32_t last_k = n_stripes - 1;
    const uint32_t m_end = (last_k / TW) * period + (last_k % TW) * kLag + S;

    // per-stripe state, live across macro steps
    uint32_t i = 0;
    bool valid = false, last_row = false;
    unsigned s_o = 0, s_e = 0, s_to = 0, s_te = 0, k_e = 0, k_te = 0, g1o = 0, g1t = 0, nongap1 = 0;
    long long srgo = 0, srge = 0, srto = 0, srte = 0, col0cost = 0;
    Cell cur = {kNeg, kNeg, kNeg, 0}, up = {kNeg, kNeg, kNeg, 0};
    long long t_next = 0;
    unsigned char* dk = dirs;          // this stripe's skewed directions / T: element (s, lane) at [s * 32 + lane]
    const long long* tk = T;
    // columns c0 .. c0+CH-1 of the column records -> this warp's shared-memory window (16-byte copies, two columns each)
    auto request_columns = [&](uint32_t c0) {
        if (VAR == 0) return;
        constexpr int per_field = CH / 2;
#pragma unroll
        for (int idx = (int)lane; idx < kColFields * per_field; idx += 32) {
            const int f = idx / per_field;
            const uint32_t c = c0 + 2u * (uint32_t)(idx % per_field);
            if (c <= WC) cp_async16(&ring[f][c & (kRing - 1)], col + (size_t)f * cstride + c);
        }
    };

    for (uint32_t m = 0; m < (TW == 1 ? rounds * S : m_end); ++m) {
        // my stripes start at r*period + team_warp*kLag, r = 0, 1, ...; `off` = local macro step inside the stripe
        const int rel = (int)m - (int)(team_warp * kLag);
        const uint32_t r = rel >= 0 ? (uint32_t)rel / period : 0;
        const int off = rel >= 0 ? (int)((uint32_t)rel - r * period) : -1;
        const uint32_t k = r * TW + team_warp;
        const bool mine = off >= 0 && off < (int)S && k < n_stripes;        // warp-uniform
        if (mine) {
            if (off == 0) {
                // ---- new stripe: row-side constants into registers
                i = k * 32 + 1 + lane;
                valid = i <= WR;
                last_row = i == WR;
                s_o = s_e = s_to = s_te = k_e = k_te = g1o = g1t = nongap1 = 0;
                srgo = srge = srto = srte = col0cost = 0;
                if (valid) {
                    if (VAR == 2) {
                        const int* rc = M.CR + (size_t)i * 32;
                        int a, b, c, d, e, f;
                        solve_gaps(M.CR, i, WR, M.nR, a, b, c, d, e, f);
                        s_o = (unsigned)a; s_e = (unsigned)b; s_to = (unsigned)c; s_te = (unsigned)d; k_e = (unsigned)e; k_te = (unsigned)f;
                        g1o = (unsigned)rc[kGO]; g1t = (unsigned)rc[kTO];
                        for (int q = 0; q < 24; ++q) nongap1 += (unsigned)rc[q];
                        const long long* sr = M.SR + (size_t)i * 32;
                        srgo = sr[kGO]; srge = sr[kGE]; srto = sr[kTO]; srte = sr[kTE];
                        col0cost = (i == 1 ? srto : srte) * M.nC;
                    } else if (VAR == 1) col0cost = (i == 1 ? to : te) * M.nC;
                    else col0cost = i == 1 ? to : te;
                }
                cur = Cell{kNeg, kNeg, kNeg, 0};
                up = Cell{kNeg, kNeg, kNeg, 0};
                t_next = 0;
                dk = dirs + (size_t)k * 32 * steps;
                tk = T + (size_t)k * 32 * steps;
                // first boundary / column-record chunk (columns 0..kChunk-1): nobody could prefetch it for us
                if (lane < kChunk && lane <= WC) cp_async_cell(&sb[0][lane], brow + lane);
                request_columns(0);
                cp_async_commit();
            }
            // the chunk of this macro step was requested one macro step ago (or just above)
            cp_async_wait_all();
            __syncwarp();
            {   // request the next one: columns (off+1)*kChunk ... have been parked (see kLag)
                const uint32_t nb = (uint32_t)(off + 1) * kChunk;
                if (lane < kChunk && nb + lane <= WC) cp_async_cell(&sb[(off + 1) & 1][lane], brow + nb + lane);
                request_columns(nb);
                cp_async_commit();
            }
            const Cell* chunk = sb[off & 1];
            const uint32_t s_begin = (uint32_t)off * kChunk;
            if (off == 0) {
                // Column 0 of the stripe (profile_par.cpp:625-640) in closed form, so that the step below never sees
                // j == 0:  D = H = NEG and V(i, 0) = max(D, V)(i-1, 0) + cost_i, a running sum down the rows (D(i-1, 0)
                // is NEG below row 0).  `cur` starts as that cell, its direction byte (all-V) and, for the stripe's
                // last row, its boundary-row copy are written here -- after this stripe has read the old brow[0].
                const Cell B = chunk[0];
                long long pre = col0cost;                               // 0 in lanes past the last row
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const long long v = shfl_up_ll_by(pre, o);
                    if ((int)lane >= o) pre += v;
                }
                cur = Cell{kNeg, kNeg, last_row ? kNeg : (B.D > B.V ? B.D : B.V) + pre, 0};
                if (valid) {
                    dk[(size_t)lane * 32 + lane] = (unsigned char)(2 | 2 << 2 | 2 << 4);
                    if (lane == 31 && !last_row) store_cell(brow, cur);
                }
            }
#pragma unroll 2
            for (uint32_t u = 0; u < (uint32_t)kChunk; ++u) {
                const uint32_t s = s_begin + u;
                // The whole body is executed by all 32 lanes (results are committed under `active`), so the warp
                // never diverges around the shuffles; steps past the end of the stripe (the last chunk) commit nothing.
                const int j = (int)s - (int)lane;                        // column handled now
                const bool active = valid && j >= 1 && j <= (int)WC;
                const long long t = t_next;
                // T of the next step: one coalesced 256-byte read per warp; the lines of the step after next few are
                // pulled into L1 by two lanes
                if (s + 1 < steps) t_next = tk[(size_t)(s + 1) * 32 + lane];
                if (lane < 2 && s + kPrefetch < steps) prefetch_l1(tk + (size_t)(s + kPrefetch) * 32 + lane * 16);
                // this column's record from the shared-memory window (consecutive lanes -> consecutive slots)
                ColInfo ci = {0, 0, 0, 0, 0, 0, 0, 0};
                if (VAR != 0) {
                    const uint32_t slot = (uint32_t)j & (kRing - 1);
                    ci.cgo = ring[0][slot]; ci.cge = ring[1][slot]; ci.cto = ring[2][slot]; ci.cte = ring[3][slot];
                    ci.chg = ring[4][slot]; ci.b0 = ring[5][slot]; ci.b1 = ring[6][slot];
                    if (VAR == 2) ci.b2 = ring[7][slot];
                }
                // (i-1, j): the lane above computed it one step ago; lane 0 takes it from the boundary row
                Cell U;
                U.D = shfl_up_ll(cur.D); U.H = shfl_up_ll(cur.H); U.V = shfl_up_ll(cur.V);
                {
                    const Cell B = chunk[u];                             // broadcast read; only lane 0 keeps it
                    const bool take = lane == 0;
                    U.D = take ? B.D : U.D; U.H = take ? B.H : U.H; U.V = take ? B.V : U.V;
                }
                const Cell Pd = up;                                      // (i-1, j-1)
                up = U;

                const Cell L = cur;
                const bool three = i > 1 && j > 1;
                Cell out;
                out.pad = 0;
                int dD, dH, dV;
                if (VAR == 0) {
                    // profile_seq.cpp:86-140 (note the >= in the second D test)
                    const bool dw = (Pd.D > Pd.H) & (Pd.D > Pd.V), hw = Pd.H >= Pd.V;
                    out.D = (dw ? Pd.D : (hw ? Pd.H : Pd.V)) + t;
                    dD = dw ? 0 : (hw ? 1 : 2);
                    long long tD = L.D + (!last_row ? go : to);
                    const long long tH = L.H + (!last_row ? ge : te);
                    out.H = tD > tH ? tD : tH; dH = tD > tH ? 0 : 1;
                    tD = U.D + (j < (int)WC ? go : to);
                    const long long tV = U.V + (j < (int)WC ? ge : te);
                    out.V = tD > tV ? tD : tV; dV = tD > tV ? 0 : 2;
                } else if (VAR == 1) {
                    // profile_par.cpp:255-421
                    dD = pick3(Pd.D, Pd.H, Pd.V + ci.chg, 0, 1, 2, out.D);
                    out.D += t;
                    const long long gcH = !last_row ? ci.cgo : ci.cto;
                    long long tD = L.D + gcH;
                    const long long tH = L.H + (!last_row ? ci.cge : ci.cte);
                    dH = pick3(tD, three ? L.V + gcH : kNever, tH, 0, 2, 1, out.H);
                    tD = U.D + ci.b0;
                    const long long tV = U.V + ci.b1;
                    dV = pick3(tD, three ? U.H + ci.b0 : kNever, tV, 0, 1, 2, out.V);
                } else {
                    // profile_par.cpp:679-886
                    long long tD = Pd.D + t;
                    long long tH = Pd.H + t;
                    tH += (ci.cge - ci.cgo) * g1o + (ci.cte - ci.cto) * g1t;       // == 0 when both counts are 0
                    long long tV = Pd.V + t + ci.chg * nongap1;
                    dD = pick3(tD, tH, tV, 0, 1, 2, out.D);
                    const long long gcH = ci.cgo * s_o + ci.cge * s_e + ci.cto * s_to + ci.cte * s_te;
                    tD = L.D + gcH;
                    tH = L.H + ci.cge * k_e + ci.cte * k_te;
                    dH = pick3(tD, three ? L.V + gcH : kNever, tH, 0, 2, 1, out.H);
                    const long long gcV = srgo * ulo32(ci.b0) + srge * uhi32(ci.b0) + srto * ulo32(ci.b1) + srte * uhi32(ci.b1);
                    tD = U.D + gcV;
                    tV = U.V + srge * ulo32(ci.b2) + srte * uhi32(ci.b2);
                    dV = pick3(tD, three ? U.H + gcV : kNever, tV, 0, 1, 2, out.V);
                }
                const unsigned char db = (unsigned char)(dD | dH << 2 | dV << 4);
                // Commit as straight-line code (selects + predicated stores).  `cur` only has to be protected while the
                // lane still waits for its first column; what it holds past the last column is never read.
                const bool commit = j >= 1;
                cur.D = commit ? out.D : cur.D; cur.H = commit ? out.H : cur.H; cur.V = commit ? out.V : cur.V;
                if (active) dk[(size_t)s * 32 + lane] = db;
                if (active && lane == 31 && !last_row) store_cell(brow + j, out);   // park the stripe's last row (L2)
                if (active && last_row && j == (int)WC) { last_out[0] = out.D; last_out[1] = out.H; last_out[2] = out.V; }
            }
        }
        // hand-over point: parked columns become visible to the next stripe
        if (CL > 1) cooperative_groups::this_cluster().sync();
        else if (NW > 1) __syncthreads();
        else __syncwarp();
    }
}

// NW == 1: four independent merges per 128-thread block (one warp each).  NW > 1: one merge per block.  CL > 1: one
// merge per thread-block CLUSTER of CL blocks (the very wide merges near the root of the guide tree, where the
// reference switches to its multi-threaded ParAlign* variants): the staircase then spans NW*CL warps on CL SMs,
// the boundary row travels through L2 and the hand-over barrier is the cluster barrier.
template <int NW, int CL>
__global__ void __launch_bounds__((NW == 1 ? kDpWarps : NW) * 32) k_dp_fill(const DpParams P)
{
    constexpr int kBlockWarps = NW == 1 ? kDpWarps : NW;
    __shared__ long long sm_last[kBlockWarps][3];
    constexpr int CH = kChunk;
    __shared__ __align__(16) Cell sm_brow[kBlockWarps][2][CH];
    __shared__ unsigned char sm_tile[NW == 1 ? kDpWarps : 1][32 * 32];
    __shared__ __align__(16) long long sm_ring[kBlockWarps][kColFields][kRing];
    const uint32_t warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t cta_rank = CL > 1 ? cooperative_groups::this_cluster().block_rank() : 0;
    const uint32_t team_warp = NW == 1 ? 0 : cta_rank * NW + warp;
    const uint32_t slot = NW == 1 ? blockIdx.x * kDpWarps + warp : blockIdx.x / CL;
    if (slot >= P.n_jobs) return;
    auto team_sync = [&]() {
        if (CL > 1) cooperative_groups::this_cluster().sync();
        else if (NW == 1) __syncwarp();
        else __syncthreads();
    };
    const uint32_t jid = P.order[slot];
    const DpJobDev J = P.jobs[jid];
    const DpMeta M = P.meta[jid];

    const Scratch L(J.w1, J.w2);
    unsigned char* scratch = P.scratch + J.scratch_off;
    const long long* col = reinterpret_cast<const long long*>(scratch + L.col);
    const uint32_t cstride = (uint32_t)L.cstride;
    Cell* brow = reinterpret_cast<Cell*>(scratch + L.brow);
    unsigned char* tmp_path = scratch + L.tmp;
    long long* g_last = reinterpret_cast<long long*>(scratch + L.lastv);
    unsigned char* dirs = P.sdirs + J.t_off;
    const long long* T = P.T + J.t_off;
    const uint32_t WR = M.WR, WC = M.WC;
    const size_t steps = (size_t)WC + 32;

    // the lane that owns cell (WR, WC) leaves (D,H,V) in shared memory, or in the job's scratch for a cluster
    long long* last_out = CL > 1 ? g_last : sm_last[warp];
    if (M.var == 0) dp_stripes<0, NW, CL, CH>(P, M, T, col, cstride, brow, dirs, team_warp, last_out, sm_brow[warp], sm_ring[warp]);
    else if (M.var == 1) dp_stripes<1, NW, CL, CH>(P, M, T, col, cstride, brow, dirs, team_warp, last_out, sm_brow[warp], sm_ring[warp]);
    else dp_stripes<2, NW, CL, CH>(P, M, T, col, cstride, brow, dirs, team_warp, last_out, sm_brow[warp], sm_ring[warp]);
    __threadfence();
    team_sync();
    if (team_warp != 0) return;

    // the warp that owned the final stripe stored (D,H,V)(WR,WC)
    const uint32_t owner_warp = NW == 1 ? warp : ((WR + 31) / 32 - 1) % NW;
    long long last[3];
    if (CL > 1) { last[0] = __ldcg(g_last); last[1] = __ldcg(g_last + 1); last[2] = __ldcg(g_last + 2); }
    else { last[0] = sm_last[owner_warp][0]; last[1] = sm_last[owner_warp][1]; last[2] = sm_last[owner_warp][2]; }

    // ---- traceback (ConstructProfile, profile.cpp:727-775).  The warp fetches the 32 x 32 corner of the direction
    // matrix that ends at the current cell into shared memory (32 independent byte loads per lane instead of one
    // dependent L2 round trip per path step), lane 0 walks inside the tile, repeat.
    uint32_t n = 0;
    long long total = 0;
    {
        unsigned char* tile = sm_tile[NW == 1 ? warp : 0];
        int dir;
        if (last[0] >= last[1] && last[0] >= last[2]) { dir = 0; total = last[0]; }
        else if (last[1] > last[2]) { dir = 1; total = last[1]; }
        else { dir = 2; total = last[2]; }
        uint32_t ti = WR, tj = WC;
        while (ti || tj) {
            const uint32_t i0 = ti >= 31 ? ti - 31 : 0, j0 = tj >= 31 ? tj - 31 : 0;
            if (ti >= lane && ti - lane >= i0) {
                const uint32_t w = tj - j0 + 1;
#pragma unroll
                for (uint32_t c = 0; c < 32; ++c)
                    if (c < w) tile[lane * 32 + c] = dir_at(dirs, steps, ti - lane, j0 + c);
            }
            __syncwarp();
            if (lane == 0) {
                uint32_t ii = ti, jj = tj;
                while ((ii || jj) && ii >= i0 && jj >= j0) {
                    tmp_path[n++] = (unsigned char)dir;
                    const unsigned char b = tile[(ti - ii) * 32 + (jj - j0)];
                    if (dir == 0) {
                        dir = b & 3;
                        if (ii == 0 || jj == 0) { ii = jj = 0; break; }      // cannot happen for a valid matrix
                        --ii; --jj;
                    } else if (dir == 1) { dir = (b >> 2) & 3; --jj; }
                    else { dir = (b >> 4) & 3; --ii; }
                    if ((int)ii < (int)i0 || (int)jj < (int)j0) break;
                }
                ti = ii; tj = jj;
            }
            ti = __shfl_sync(0xffffffffu, ti, 0);
            tj = __shfl_sync(0xffffffffu, tj, 0);
            dir = __shfl_sync(0xffffffffu, dir, 0);
            __syncwarp();
        }
    }
    n = __shfl_sync(0xffffffffu, n, 0);
    __syncwarp();
    unsigned char* path = P.path + J.path_off;
    for (uint32_t k = lane; k < n; k += 32) path[k] = tmp_path[n - 1 - k];
    if (lane == 0) {
        famsa_dp_result r;
        r.total_score = total;
        r.last[0] = last[0]; r.last[1] = last[1]; r.last[2] = last[2];
        r.path_offset = J.path_off; r.dirs_offset = J.dirs_off;
        r.path_len = n; r.rows_width = WR; r.cols_width = WC;
        r.swapped = (uint8_t)M.sw; r.variant = M.bad ? (uint8_t)0xFF : (uint8_t)M.var; r.pad[0] = r.pad[1] = 0;
        P.results[jid] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

#define FB_TRY(expr)                      \
    do {                                  \
        int rc__ = (expr);                \
        if (rc__ != FAMSA_OK) return rc__; \
    } while (0)

// jobs[k].p1/p2 hold DEVICE pointers here
int dp_run_device(famsa_ctx* ctx, const famsa_dp_job* jobs, uint32_t n, const int64_t gaps[4],
                  famsa_dp_result* d_results, uint8_t* d_path, uint8_t* d_dirs, cudaStream_t st)
{
    DpState& S = ctx->dp;
    std::vector<DpJobDev> dev(n);
    unsigned long long path_off = 0, dirs_off = 0, cells = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const famsa_dp_job& j = jobs[k];
        if (j.p1.width == 0 || j.p2.width == 0 || j.p1.card == 0 || j.p2.card == 0) {
            set_error("dp job " + std::to_string(k) + ": empty profile");
            return FAMSA_E_INVALID;
        }
        if (((unsigned long long)j.p1.width + 1) * (j.p2.width + 1) > 0xffffffffull) {
            set_error("dp job " + std::to_string(k) + ": more than 2^32 matrix cells");
            return FAMSA_E_INVALID;
        }
        DpJobDev& d = dev[k];
        d.s1 = reinterpret_cast<const long long*>(j.p1.scores); d.c1 = j.p1.counters;
        d.s2 = reinterpret_cast<const long long*>(j.p2.scores); d.c2 = j.p2.counters;
        d.w1 = j.p1.width; d.card1 = j.p1.card; d.w2 = j.p2.width; d.card2 = j.p2.card;
        d.path_off = path_off; d.dirs_off = dirs_off;           // caller-visible layout: global prefix sums
        path_off += (unsigned long long)d.w1 + d.w2;
        dirs_off += ((unsigned long long)d.w1 + 1) * (d.w2 + 1);
        cells += (unsigned long long)d.w1 * d.w2;
    }
    S.last_cells = cells;
    // Sub-batches of consecutive jobs bound the device scratch (T is 8 bytes per cell): ~1 Gi cells each.
    unsigned long long max_cells = 1ull << 30;
    if (const char* e = getenv("FAMSA_DP_MAX_CELLS")) max_cells = strtoull(e, nullptr, 10);     // development knob
    uint32_t team_min = kDpTeamMinWidth;
    if (const char* e = getenv("FAMSA_DP_TEAM_MIN")) team_min = (uint32_t)atoi(e);               // development knob
    int nw_forced = 0;
    if (const char* e = getenv("FAMSA_DP_TEAM_WARPS")) nw_forced = atoi(e);                      // development knob
    uint32_t cluster_min = kDpClusterMinWidth;
    if (const char* e = getenv("FAMSA_DP_CLUSTER_MIN")) cluster_min = (uint32_t)atoi(e);         // development knob

    FB_CUDA(cudaEventRecord(ctx->ev[0], st));
    FB_CUDA(cudaEventRecord(ctx->ev[1], st));
    for (uint32_t j0 = 0; j0 < n;) {
        // [j0, j1): as many consecutive jobs as fit
        uint32_t j1 = j0;
        unsigned long long mat_sum = 0, scratch_off = 0, t_off = 0;
        std::vector<unsigned long long> tblock(1, 0), t2block(1, 0);
        while (j1 < n) {
            const unsigned long long mat = ((unsigned long long)dev[j1].w1 + 1) * (dev[j1].w2 + 1);
            if (j1 > j0 && mat_sum + mat > max_cells) break;
            dev[j1].scratch_off = scratch_off;
            dev[j1].t_off = t_off;                              // skewed T / directions are per sub-batch
            scratch_off += Scratch(dev[j1].w1, dev[j1].w2).total;
            t_off += skew_elems(dev[j1].w1, dev[j1].w2);
            mat_sum += mat;
            tblock.push_back(tblock.back() + (mat + kTThreads * kTCellsPerThread - 1) / (kTThreads * kTCellsPerThread));
            t2block.push_back(t2block.back() + t_blocks(dev[j1].w1, dev[j1].w2));
            ++j1;
        }
        const uint32_t m = j1 - j0;
        if (tblock[m] > 0x7fffffffull || t2block[m] > 0x7fffffffull) { set_error("dp sub-batch too large for one launch"); return FAMSA_E_INVALID; }
        // merges whose shorter side spans several 32-row stripes get a whole block (a team of warps pipelined over
        // the stripes); the rest run one warp per merge.  Both groups cost-descending.
        // class 2: shorter side > cluster_min -> a cluster of blocks; class 1: > team_min -> one block; class 0: one warp
        // When a level holds only a handful of block-sized merges (the top of the guide tree) the GPU would sit idle:
        // give every merge with more than 8 stripes a cluster then.
        uint32_t n_teamable = 0;
        for (uint32_t a = j0; a < j1; ++a) n_teamable += std::min(dev[a].w1, dev[a].w2) > team_min;
        uint32_t cl_min = cluster_min;
        if (n_teamable * kDpCluster <= 2u * (uint32_t)ctx->sm_count) cl_min = std::min(cluster_min, std::max(team_min, 256u));
        auto cls = [&](uint32_t a) { const uint32_t w = std::min(dev[a].w1, dev[a].w2); return w > cl_min ? 2 : (w > team_min ? 1 : 0); };
        std::vector<uint32_t> order(m);
        std::iota(order.begin(), order.end(), j0);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            if (cls(a) != cls(b)) return cls(a) > cls(b);
            return (unsigned long long)dev[a].w1 * dev[a].w2 > (unsigned long long)dev[b].w1 * dev[b].w2;
        });
        uint32_t n_huge = 0, n_big = 0;
        while (n_huge < m && cls(order[n_huge]) == 2) ++n_huge;
        n_big = n_huge;
        while (n_big < m && cls(order[n_big]) == 1) ++n_big;

        FB_TRY(S.d_jobs.reserve(sizeof(DpJobDev) * n));
        FB_TRY(S.d_meta.reserve(sizeof(DpMeta) * n));
        FB_TRY(S.d_order.reserve(sizeof(uint32_t) * m));
        FB_TRY(S.d_tblock.reserve(sizeof(unsigned long long) * 2 * (m + 1)));
        FB_TRY(S.d_scratch.reserve(std::max<unsigned long long>(scratch_off, 64)));
        FB_TRY(S.d_T.reserve(std::max<unsigned long long>(t_off * 8, 64)));
        FB_TRY(S.d_dirs.reserve(std::max<unsigned long long>(t_off, 64)));
        FB_CUDA(cudaMemcpyAsync(S.d_jobs.as<DpJobDev>() + j0, dev.data() + j0, sizeof(DpJobDev) * m, cudaMemcpyHostToDevice, st));
        FB_CUDA(cudaMemcpyAsync(S.d_order.p, order.data(), sizeof(uint32_t) * m, cudaMemcpyHostToDevice, st));
        FB_CUDA(cudaMemcpyAsync(S.d_tblock.p, tblock.data(), sizeof(unsigned long long) * (m + 1), cudaMemcpyHostToDevice, st));
        FB_CUDA(cudaMemcpyAsync(S.d_tblock.as<unsigned long long>() + (m + 1), t2block.data(), sizeof(unsigned long long) * (m + 1),
                                cudaMemcpyHostToDevice, st));
        DpParams P{};
        P.jobs = S.d_jobs.as<DpJobDev>();
        P.meta = S.d_meta.as<DpMeta>();
        P.order = S.d_order.as<uint32_t>();
        P.n_jobs = m;
        P.job_base = j0;
        P.go = gaps[0]; P.ge = gaps[1]; P.to = gaps[2]; P.te = gaps[3];
        P.dirs = d_dirs;
        P.sdirs = S.d_dirs.as<uint8_t>();
        P.path = d_path;
        P.scratch = S.d_scratch.as<uint8_t>();
        P.T = S.d_T.as<long long>();
        P.tblock = S.d_tblock.as<unsigned long long>();
        P.t2block = P.tblock + (m + 1);
        P.results = d_results;
        k_dp_prep<<<m, kPrepThreads, 0, st>>>(P);
        FB_CUDA(cudaGetLastError());
        k_dp_t<<<(unsigned)t2block[m], kTCols, 0, st>>>(P);
        FB_CUDA(cudaGetLastError());
        ctx->launches += 2;
        // `order`: cluster jobs, then block jobs, then warp jobs (see the sort above)
        if (n_huge) {
            DpParams Q = P;
            Q.n_jobs = n_huge;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(n_huge * kDpCluster);
            cfg.blockDim = dim3(kDpTeamWarps * 32);
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = kDpCluster;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            FB_CUDA(cudaLaunchKernelEx(&cfg, k_dp_fill<kDpTeamWarps, kDpCluster>, Q));
            ctx->launches++;
        }
        if (n_big > n_huge) {
            DpParams Q = P;
            Q.order = P.order + n_huge;
            Q.n_jobs = n_big - n_huge;
            // Team size by how many merges there are: a team's ramp-up / ramp-down (stripes start 6 macro steps apart)
            // idles a third of an 8-warp team on ~14-stripe merges, so when the level is large enough to fill the SMs with
            // smaller teams (16 warps per SM either way) those waste less -- 4 warps from 4 merges per SM on, 2 from 8.
            int nw = kDpTeamWarps;
            if (Q.n_jobs >= 8u * (uint32_t)ctx->sm_count) nw = 2;
            else if (Q.n_jobs >= 4u * (uint32_t)ctx->sm_count) nw = 4;
            if (nw_forced) nw = nw_forced;
            switch (nw) {
            case 2: k_dp_fill<2, 1><<<Q.n_jobs, 2 * 32, 0, st>>>(Q); break;
            case 4: k_dp_fill<4, 1><<<Q.n_jobs, 4 * 32, 0, st>>>(Q); break;
            default: k_dp_fill<kDpTeamWarps, 1><<<Q.n_jobs, kDpTeamWarps * 32, 0, st>>>(Q); break;
            }
            FB_CUDA(cudaGetLastError());
            ctx->launches++;
        }
        if (m > n_big) {
            DpParams Q = P;
            Q.order = P.order + n_big;
            Q.n_jobs = m - n_big;
            k_dp_fill<1, 1><<<(Q.n_jobs + kDpWarps - 1) / kDpWarps, kDpWarps * 32, 0, st>>>(Q);
            FB_CUDA(cudaGetLastError());
            ctx->launches++;
        }
        if (d_dirs) {                                            // caller wants CDPMatrix bytes: un-skew
            k_dp_unskew<<<(unsigned)tblock[m], kTThreads, 0, st>>>(P);
            FB_CUDA(cudaGetLastError());
            ctx->launches++;
        }
        // the next sub-batch reuses the scratch: the reserve() calls above may also free+reallocate, and
        // cudaFree synchronises, so nothing is released while kernels still read it
        j0 = j1;
    }
    FB_CUDA(cudaEventRecord(ctx->ev[2], st));
    FB_CUDA(cudaEventRecord(ctx->ev[3], st));
    return FAMSA_OK;
}

// variant 0xFF: k_dp_prep found a negative count in the tables (see ulo32) -- fail loudly rather than return a
// result that could differ from the reference's signed arithmetic
int dp_check_results(const famsa_dp_result* results, uint32_t n)
{
    for (uint32_t k = 0; k < n; ++k)
        if (results[k].variant == 0xFF) {
            set_error("dp job " + std::to_string(k) + ": a profile holds negative residue / gap counts");
            return FAMSA_E_INVALID;
        }
    return FAMSA_OK;
}

int dp_run_host(famsa_ctx* ctx, const famsa_dp_job* jobs, uint32_t n, const int64_t gaps[4], famsa_dp_result* results,
                uint8_t* path_buf, uint8_t* dirs_buf)
{
    DpState& S = ctx->dp;
    cudaStream_t st = ctx->stream;
    // pack every table into one staging buffer -> one H2D
    unsigned long long bytes = 0, path_total = 0, dirs_total = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const famsa_dp_job& j = jobs[k];
        if (!j.p1.scores || !j.p1.counters || !j.p2.scores || !j.p2.counters) {
            set_error("dp job " + std::to_string(k) + ": NULL table");
            return FAMSA_E_INVALID;
        }
        bytes += ((unsigned long long)j.p1.width + 1 + j.p2.width + 1) * 32 * (8 + 4);
        path_total += (unsigned long long)j.p1.width + j.p2.width;
        dirs_total += ((unsigned long long)j.p1.width + 1) * (j.p2.width + 1);
    }
    // pinned staging buffer; tables are packed and shipped in ~8 MB slices so that packing slice k+1 overlaps the
    // H2D of slice k
    if (bytes > S.h_pinned_cap) {
        if (S.h_pinned) cudaFreeHost(S.h_pinned);
        S.h_pinned = nullptr;
        S.h_pinned_cap = 0;
        FB_CUDA(cudaHostAlloc(&S.h_pinned, bytes + bytes / 4 + 4096, cudaHostAllocDefault));
        S.h_pinned_cap = bytes + bytes / 4 + 4096;
    }
    FB_TRY(S.d_tables.reserve(std::max<unsigned long long>(bytes, 64)));
    std::vector<famsa_dp_job> dj(jobs, jobs + n);
    unsigned long long at = 0, shipped = 0;
    uint8_t* hb = static_cast<uint8_t*>(S.h_pinned);
    uint8_t* db = S.d_tables.as<uint8_t>();
    // device addresses first (cheap), then the copies: the staging buffer is filled by a few host threads, each
    // shipping its own contiguous slice as soon as it is packed
    std::vector<unsigned long long> job_at(n + 1, 0);
    for (uint32_t k = 0; k < n; ++k) {      // every table size is a multiple of 128 bytes, so alignment is kept
        famsa_dp_job& j = dj[k];
        const size_t s1 = ((size_t)jobs[k].p1.width + 1) * 32 * 8, s2 = ((size_t)jobs[k].p2.width + 1) * 32 * 8;
        j.p1.scores = reinterpret_cast<const int64_t*>(db + at);
        j.p2.scores = reinterpret_cast<const int64_t*>(db + at + s1);
        j.p1.counters = reinterpret_cast<const int32_t*>(db + at + s1 + s2);
        j.p2.counters = reinterpret_cast<const int32_t*>(db + at + s1 + s2 + s1 / 2);
        at += (s1 + s2) * 3 / 2;
        job_at[k + 1] = at;
    }
    {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        const unsigned n_thr = bytes > (16u << 20) ? std::min(8u, hw) : 1;
        std::vector<std::thread> workers;
        std::vector<cudaError_t> errs(n_thr, cudaSuccess);
        for (unsigned t = 0; t < n_thr; ++t)
            workers.emplace_back([&, t] {
                cudaSetDevice(ctx->device);
                const uint32_t k0 = (uint32_t)((unsigned long long)n * t / n_thr), k1 = (uint32_t)((unsigned long long)n * (t + 1) / n_thr);
                unsigned long long sent = job_at[k0];
                for (uint32_t k = k0; k < k1; ++k) {
                    const size_t s1 = ((size_t)jobs[k].p1.width + 1) * 32 * 8, s2 = ((size_t)jobs[k].p2.width + 1) * 32 * 8;
                    uint8_t* h = hb + job_at[k];
                    memcpy(h, jobs[k].p1.scores, s1);
                    memcpy(h + s1, jobs[k].p2.scores, s2);
                    memcpy(h + s1 + s2, jobs[k].p1.counters, s1 / 2);
                    memcpy(h + s1 + s2 + s1 / 2, jobs[k].p2.counters, s2 / 2);
                    if (job_at[k + 1] - sent >= (8u << 20) || k + 1 == k1) {
                        cudaError_t e = cudaMemcpyAsync(db + sent, hb + sent, job_at[k + 1] - sent, cudaMemcpyHostToDevice, st);
                        if (e != cudaSuccess) errs[t] = e;
                        sent = job_at[k + 1];
                    }
                }
            });
        for (auto& w : workers) w.join();
        for (cudaError_t e : errs) FB_CUDA(e);
    }
    (void)shipped;
    FB_TRY(S.d_results.reserve(sizeof(famsa_dp_result) * std::max(1u, n)));
    FB_TRY(S.d_path.reserve(std::max<unsigned long long>(path_total, 64)));
    uint8_t* d_dirs = nullptr;
    if (dirs_buf) {
        FB_TRY(S.d_dirs_out.reserve(std::max<unsigned long long>(dirs_total, 64)));   // row-major copy for the caller
        d_dirs = S.d_dirs_out.as<uint8_t>();
    }
    FB_TRY(dp_run_device(ctx, dj.data(), n, gaps, S.d_results.as<famsa_dp_result>(), S.d_path.as<uint8_t>(), d_dirs, st));
    if (n) FB_CUDA(cudaMemcpyAsync(results, S.d_results.p, sizeof(famsa_dp_result) * n, cudaMemcpyDeviceToHost, st));
    if (path_total) FB_CUDA(cudaMemcpyAsync(path_buf, S.d_path.p, path_total, cudaMemcpyDeviceToHost, st));
    if (dirs_buf && dirs_total) FB_CUDA(cudaMemcpyAsync(dirs_buf, d_dirs, dirs_total, cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    return dp_check_results(results, n);
}

} // namespace fb
