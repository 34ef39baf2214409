// HP-2: profile-profile / sequence-profile / sequence-sequence affine-gap DP with position-specific
// gap costs, many independent merges per launch, plus on-device traceback.
//
// Replaces CProfile::Align and its cell loops (reference src/core/profile.cpp:244-305,
// profile_seq.cpp:24-892, profile_par.cpp:26-903), DP_SolveGapsProblemWhenStarting/Continuing
// (profile.cpp:1223-1315) and the traceback of ConstructProfile (profile.cpp:727-782).
// Semantics follow SURVEY.md Appendix B; arithmetic is int64 with NEG = -(1<<62) unsaturated.
//
// Design: one warp per merge.  The DP matrix is swept in stripes of 32 rows; inside a stripe
// lane L owns row i0+L and at step s computes column s-L (an anti-diagonal wavefront), receiving
// the (D,H,V) of the cell above through warp shuffles and keeping its left neighbour in registers.
// The last row of a stripe is parked in a per-merge boundary row (global, L1/L2 resident) and
// feeds lane 0 of the next stripe.  Column-side constants (gap-correction counts, the profile-2
// score column) are read through L1; row-side constants live in registers / shared memory.
// After the fill the same warp walks the direction matrix back and emits the path.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "ctx.h"

namespace fb {

constexpr int kGO = 25, kGE = 26, kTE = 27, kTO = 28;   // GAP_OPEN, GAP_EXT, GAP_TERM_EXT, GAP_TERM_OPEN (defs.h:62-66)
constexpr long long kNeg = -(1ll << 62);
constexpr int kDpWarps = 4;
constexpr int kDpTeamWarps = 8;          // warps cooperating on one large merge
constexpr uint32_t kDpTeamMinWidth = 96;  // min(w1, w2) above which a merge gets a team

struct DpJobDev {
    const long long* s1; const int* c1;
    const long long* s2; const int* c2;
    uint32_t w1, card1, w2, card2;
    unsigned long long path_off, dirs_off, scratch_off;
};

struct ColInfo {           // per column j of the column profile (64 bytes)
    int s_o, s_e, s_to, s_te, k_e, k_te;   // DP_SolveGapsProblemWhenStarting / Continuing
    int sym;                               // residue if the column profile is a single sequence
    int pad;
    long long chg2;                        // cnt[GO]*(ge-go) + cnt[TO]*(te-to)
    long long gcv1, contv1;                // SeqProf: scalar-gap V costs (profile_par.cpp:204-211)
    long long pad2[1];
};
static_assert(sizeof(ColInfo) == 64, "ColInfo layout");

struct Cell { long long D, H, V; };

__device__ __forceinline__ long long shfl_up_ll(long long v)
{
    int lo = (int)(unsigned long long)v, hi = (int)((unsigned long long)v >> 32);
    lo = __shfl_up_sync(0xffffffffu, lo, 1);
    hi = __shfl_up_sync(0xffffffffu, hi, 1);
    return (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo);
}

// boundary-row traffic goes through L2 (.cg): written by lane 31 of one stripe, read by lane 0 of the next
__device__ __forceinline__ Cell load_cell(const Cell* p)
{
    const long long* q = reinterpret_cast<const long long*>(p);
    Cell c; c.D = __ldcg(q); c.H = __ldcg(q + 1); c.V = __ldcg(q + 2);
    return c;
}
__device__ __forceinline__ void store_cell(Cell* p, const Cell& c)
{
    long long* q = reinterpret_cast<long long*>(p);
    __stcg(q, c.D); __stcg(q + 1, c.H); __stcg(q + 2, c.V);
}

// a if a>b && a>c; else b if b>c; else c  (strict comparisons, fixed priority)
__device__ __forceinline__ int pick3(long long a, long long b, long long c, int da, int db, int dc, long long& out)
{
    if (a > b && a > c) { out = a; return da; }
    if (b > c) { out = b; return db; }
    out = c; return dc;
}

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
    const uint32_t* order;        // launch slot -> job index (cost-descending)
    uint32_t n_jobs;
    long long go, ge, to, te;
    unsigned char* dirs;          // all direction matrices
    unsigned char* path;          // all paths (forward order)
    unsigned char* scratch;       // per job: ColInfo[wmax+1], Cell brow[wmax+1], tmp path[w1+w2]
    famsa_dp_result* results;
};

// One team = NW warps working on one merge.  Stripe k (rows 32k+1 .. 32k+32) belongs to warp k % NW;
// consecutive stripes run as a staircase: stripe k+1 may touch column j only after stripe k has parked
// its last row's cell (., j) in `brow`.  prog[w] is warp w's monotonically increasing count of parked
// columns ((round * (WC+1)) + columns of the current stripe), polled by the warp that owns the next stripe.
template <int VAR, int NW>
__device__ __forceinline__ void dp_stripes(const DpParams& P, const long long* __restrict__ SRs, const int* __restrict__ CR,
                                           const long long* __restrict__ SCs, uint32_t WR, uint32_t WC, int nR, int nC,
                                           const ColInfo* __restrict__ col, Cell* __restrict__ brow,
                                           unsigned char* __restrict__ dirs, int* __restrict__ nz_k, int* __restrict__ nz_c,
                                           volatile unsigned* prog, uint32_t team_warp, long long* last_out)
{
    const uint32_t lane = threadIdx.x & 31;
    const size_t ld = (size_t)WC + 1;
    const long long go = P.go, ge = P.ge, to = P.to, te = P.te;
    const uint32_t n_stripes = (WR + 31) / 32;

    for (uint32_t k = team_warp; k < n_stripes; k += NW) {
        const uint32_t i = k * 32 + 1 + lane;
        const bool valid = i <= WR;
        const bool last_row = i == WR;
        // what the producer (stripe k-1, warp (k-1) % NW) must have published before column j may be read
        const uint32_t prod_warp = (k + NW - 1) % NW;
        const unsigned prod_base = k ? ((k - 1) / NW) * (WC + 1) : 0;
        const unsigned my_base = (k / NW) * (WC + 1);
        unsigned avail = k ? 0 : WC + 1;                 // columns of the producer known to be parked
        // ---- row-side constants
        int symR = 22, s_o = 0, s_e = 0, s_to = 0, s_te = 0, k_e = 0, k_te = 0, g1o = 0, g1t = 0, nzn = 0;
        long long nongap1 = 0, srgo = 0, srge = 0, srto = 0, srte = 0, col0cost = 0;
        if (valid) {
            const int* rc = CR + (size_t)i * 32;
            if (VAR != 2) symR = seq_symbol(CR, i);
            if (VAR == 2) {
                solve_gaps(CR, i, WR, nR, s_o, s_e, s_to, s_te, k_e, k_te);
                g1o = rc[kGO]; g1t = rc[kTO];
                for (int q = 0; q < 30; ++q) {
                    const int c = rc[q];
                    if (c) { nz_k[nzn * 32 + lane] = q; nz_c[nzn * 32 + lane] = c; ++nzn; if (q < 24) nongap1 += c; }
                }
                const long long* sr = SRs + (size_t)i * 32;
                srgo = sr[kGO]; srge = sr[kGE]; srto = sr[kTO]; srte = sr[kTE];
                col0cost = (i == 1 ? srto : srte) * nC;
            } else if (VAR == 1) col0cost = (i == 1 ? to : te) * nC;
            else col0cost = i == 1 ? to : te;
        }
        __syncwarp();

        auto wait_for = [&](uint32_t j) {               // lane 0 only: producer has parked column j
            if (NW > 1 && j >= avail) {
                unsigned v;
                do { v = prog[prod_warp] - prod_base; } while ((int)v < 0 || v <= j);
                avail = v;
                __threadfence_block();
            }
        };

        Cell cur = {kNeg, kNeg, kNeg};      // own cell of the previous step (left neighbour)
        Cell up = {kNeg, kNeg, kNeg};       // cell above of the previous step (becomes the diagonal)
        Cell nxt = {kNeg, kNeg, kNeg};      // lane 0: boundary-row cell for the next step (L2 prefetch)
        if (lane == 0) { wait_for(0); nxt = load_cell(brow); }
        unsigned char* drow = dirs + (size_t)i * ld;
        const uint32_t steps = WC + 1 + 31;
        for (uint32_t s = 0; s < steps; ++s) {
            const int j = (int)s - (int)lane;               // column handled now
            // (i-1, j): lane above computed it one step ago; lane 0 reads the boundary row
            Cell U;
            U.D = shfl_up_ll(cur.D); U.H = shfl_up_ll(cur.H); U.V = shfl_up_ll(cur.V);
            if (lane == 0) {
                U = nxt;
                if (s + 1 <= WC) { wait_for(s + 1); nxt = load_cell(brow + s + 1); }
            }
            const Cell Pd = up;                              // (i-1, j-1)
            up = U;
            if (!valid || j < 0 || j > (int)WC) continue;
            Cell out;
            unsigned char db;
            if (j == 0) {
                out.D = kNeg; out.H = kNeg;
                out.V = last_row ? kNeg : (U.D > U.V ? U.D : U.V) + col0cost;
                db = 2 | 2 << 2 | 2 << 4;
            } else {
                const Cell L = cur;
                const bool three = i > 1 && j > 1;
                const long long* sc = SCs + (size_t)j * 32;
                const ColInfo ci = col[j];
                int dD, dH, dV;
                if (VAR == 0) {
                    const long long sc_ = sc[symR];
                    if (Pd.D > Pd.H && Pd.D > Pd.V) { out.D = Pd.D + sc_; dD = 0; }
                    else if (Pd.H >= Pd.V) { out.D = Pd.H + sc_; dD = 1; }
                    else { out.D = Pd.V + sc_; dD = 2; }
                    long long tD = L.D + (!last_row ? go : to), tH = L.H + (!last_row ? ge : te);
                    if (tD > tH) { out.H = tD; dH = 0; } else { out.H = tH; dH = 1; }
                    tD = U.D + (j < (int)WC ? go : to);
                    const long long tV = U.V + (j < (int)WC ? ge : te);
                    if (tD > tV) { out.V = tD; dV = 0; } else { out.V = tV; dV = 2; }
                } else if (VAR == 1) {
                    const long long t = sc[symR];
                    dD = pick3(Pd.D, Pd.H, Pd.V + ci.chg2, 0, 1, 2, out.D);
                    out.D += t;
                    const long long gcH = !last_row ? sc[kGO] : sc[kTO];
                    long long tD = L.D + gcH;
                    const long long tH = L.H + (!last_row ? sc[kGE] : sc[kTE]);
                    if (three) dH = pick3(tD, L.V + gcH, tH, 0, 2, 1, out.H);
                    else if (tD > tH) { out.H = tD; dH = 0; } else { out.H = tH; dH = 1; }
                    tD = U.D + ci.gcv1;
                    const long long tV = U.V + ci.contv1;
                    if (three) dV = pick3(tD, U.H + ci.gcv1, tV, 0, 1, 2, out.V);
                    else if (tD > tV) { out.V = tD; dV = 0; } else { out.V = tV; dV = 2; }
                } else {
                    long long t = 0;
                    for (int q = 0; q < nzn; ++q) t += (long long)nz_c[q * 32 + lane] * sc[nz_k[q * 32 + lane]];
                    const long long cgo = sc[kGO], cge = sc[kGE], cte = sc[kTE], cto = sc[kTO];
                    long long tD = Pd.D + t;
                    long long tH = Pd.H + t;
                    if (g1o || g1t) tH += (long long)g1o * (cge - cgo) + (long long)g1t * (cte - cto);
                    long long tV = Pd.V + t + ci.chg2 * nongap1;
                    dD = pick3(tD, tH, tV, 0, 1, 2, out.D);
                    const long long gcH = cgo * s_o + cge * s_e + cto * s_to + cte * s_te;
                    tD = L.D + gcH;
                    tH = L.H + cge * k_e + cte * k_te;
                    if (three) dH = pick3(tD, L.V + gcH, tH, 0, 2, 1, out.H);
                    else if (tD > tH) { out.H = tD; dH = 0; } else { out.H = tH; dH = 1; }
                    const long long gcV = srgo * ci.s_o + srge * ci.s_e + srto * ci.s_to + srte * ci.s_te;
                    tD = U.D + gcV;
                    tV = U.V + srge * ci.k_e + srte * ci.k_te;
                    if (three) dV = pick3(tD, U.H + gcV, tV, 0, 1, 2, out.V);
                    else if (tD > tV) { out.V = tD; dV = 0; } else { out.V = tV; dV = 2; }
                }
                db = (unsigned char)(dD | dH << 2 | dV << 4);
            }
            drow[j] = db;
            cur = out;
            if (last_row) {
                if (j == (int)WC) { last_out[0] = out.D; last_out[1] = out.H; last_out[2] = out.V; }
            } else if (lane == 31) {
                // park the stripe's last row for the next stripe; publish every 4th column
                store_cell(brow + j, out);
                if (NW > 1 && ((j & 3) == 3 || j == (int)WC)) {
                    __threadfence_block();
                    prog[team_warp] = my_base + (unsigned)j + 1;
                }
            }
        }
        __syncwarp();
    }
}

// NW == 1: four independent merges per 128-thread block (one warp each).  NW > 1: one merge per block.
template <int NW>
__global__ void __launch_bounds__((NW == 1 ? kDpWarps : NW) * 32) k_dp_align(const DpParams P)
{
    constexpr int kBlockWarps = NW == 1 ? kDpWarps : NW;
    extern __shared__ int sm_dyn[];                         // [kBlockWarps][2][30*32]: nz symbol ids, nz counts
    int (*sm_nz_k)[30 * 32] = reinterpret_cast<int (*)[30 * 32]>(sm_dyn);
    int (*sm_nz_c)[30 * 32] = reinterpret_cast<int (*)[30 * 32]>(sm_dyn + kBlockWarps * 30 * 32);
    __shared__ unsigned sm_prog[kBlockWarps];
    __shared__ unsigned long long sm_nz[2];
    __shared__ long long sm_last[kBlockWarps][3];
    const uint32_t warp = threadIdx.x / 32, lane = threadIdx.x % 32;
    const uint32_t team_warp = NW == 1 ? 0 : warp;
    const uint32_t tid = NW == 1 ? lane : threadIdx.x;          // index inside the team
    constexpr uint32_t kTeam = NW * 32;
    const uint32_t slot = NW == 1 ? blockIdx.x * kDpWarps + warp : blockIdx.x;
    if (slot >= P.n_jobs) return;
    auto team_sync = [&]() { if (NW == 1) __syncwarp(); else __syncthreads(); };
    const uint32_t jid = P.order[slot];
    const DpJobDev J = P.jobs[jid];
    if (threadIdx.x < kBlockWarps) sm_prog[threadIdx.x] = 0;
    if (threadIdx.x < 2) sm_nz[threadIdx.x] = 0;
    if (NW > 1) __syncthreads();

    // ---- variant and orientation (CProfile::Align, profile.cpp:254-304)
    int var, sw = 0;
    if (J.card1 == 1 && J.card2 == 1) var = 0;
    else if (J.card1 == 1) var = 1;
    else if (J.card2 == 1) { var = 1; sw = 1; }
    else {
        var = 2;
        unsigned long long nz1 = 0, nz2 = 0;
        for (size_t k = tid; k < ((size_t)J.w1 + 1) * 32; k += kTeam) nz1 += J.c1[k] != 0;
        for (size_t k = tid; k < ((size_t)J.w2 + 1) * 32; k += kTeam) nz2 += J.c2[k] != 0;
        for (int o = 16; o; o >>= 1) {
            nz1 += __shfl_xor_sync(0xffffffffu, nz1, o);
            nz2 += __shfl_xor_sync(0xffffffffu, nz2, o);
        }
        if (NW > 1) {
            if (lane == 0) { atomicAdd(&sm_nz[0], nz1); atomicAdd(&sm_nz[1], nz2); }
            __syncthreads();
            nz1 = sm_nz[0]; nz2 = sm_nz[1];
        }
        if (!(nz1 * (unsigned long long)J.w2 < nz2 * (unsigned long long)J.w1)) sw = 1;
    }
    const long long* SR = sw ? J.s2 : J.s1;  const int* CR = sw ? J.c2 : J.c1;
    const long long* SC = sw ? J.s1 : J.s2;  const int* CC = sw ? J.c1 : J.c2;
    const uint32_t WR = sw ? J.w2 : J.w1, WC = sw ? J.w1 : J.w2;
    const int nR = (int)(sw ? J.card2 : J.card1), nC = (int)(sw ? J.card1 : J.card2);
    const uint32_t wmax = J.w1 > J.w2 ? J.w1 : J.w2;

    unsigned char* scratch = P.scratch + J.scratch_off;
    ColInfo* col = reinterpret_cast<ColInfo*>(scratch);
    Cell* brow = reinterpret_cast<Cell*>(scratch + sizeof(ColInfo) * ((size_t)wmax + 1));
    unsigned char* tmp_path = scratch + (sizeof(ColInfo) + sizeof(Cell)) * ((size_t)wmax + 1);
    unsigned char* dirs = P.dirs + J.dirs_off;
    const size_t ld = (size_t)WC + 1;
    const long long go = P.go, ge = P.ge, to = P.to, te = P.te;

    // ---- column-side constants and row 0
    for (uint32_t j = tid; j <= WC; j += kTeam) {
        ColInfo ci;
        ci.pad = 0; ci.pad2[0] = 0;
        if (j >= 1) {
            solve_gaps(CC, j, WC, nC, ci.s_o, ci.s_e, ci.s_to, ci.s_te, ci.k_e, ci.k_te);
            const int* cc = CC + (size_t)j * 32;
            ci.chg2 = (long long)cc[kGO] * (ge - go) + (long long)cc[kTO] * (te - to);
            ci.sym = var == 0 ? seq_symbol(CC, j) : 22;
            ci.gcv1 = go * ci.s_o + ge * ci.s_e + to * ci.s_to + te * ci.s_te;
            ci.contv1 = ge * ci.k_e + te * ci.k_te;
        } else {
            ci.s_o = ci.s_e = ci.s_to = ci.s_te = ci.k_e = ci.k_te = 0; ci.sym = 22; ci.chg2 = 0; ci.gcv1 = 0; ci.contv1 = 0;
        }
        col[j] = ci;
        dirs[j] = j == 0 ? 0 : (unsigned char)(1 | 1 << 2 | 1 << 4);
    }
    if (tid == 0) {
        store_cell(brow, Cell{0, kNeg, kNeg});
        long long h = 0;
        for (uint32_t j = 1; j <= WC; ++j) {
            const long long* sc = SC + (size_t)j * 32;
            if (var == 0) h = j == 1 ? to : h + te;            // max(H, D=NEG) + te
            else if (var == 1) h = j == 1 ? sc[kTO] : h + sc[kTE];
            else h = j == 1 ? sc[kTO] * nR : h + sc[kTE] * nR;
            store_cell(brow + j, Cell{kNeg, j == WC ? kNeg : h, kNeg});
        }
    }
    __threadfence_block();
    team_sync();

    long long* last_out = sm_last[warp];
    if (var == 0) dp_stripes<0, NW>(P, SR, CR, SC, WR, WC, nR, nC, col, brow, dirs, sm_nz_k[warp], sm_nz_c[warp], sm_prog, team_warp, last_out);
    else if (var == 1) dp_stripes<1, NW>(P, SR, CR, SC, WR, WC, nR, nC, col, brow, dirs, sm_nz_k[warp], sm_nz_c[warp], sm_prog, team_warp, last_out);
    else dp_stripes<2, NW>(P, SR, CR, SC, WR, WC, nR, nC, col, brow, dirs, sm_nz_k[warp], sm_nz_c[warp], sm_prog, team_warp, last_out);
    __threadfence_block();
    team_sync();
    if (team_warp != 0) return;

    // the warp that owned the final stripe stored (D,H,V)(WR,WC)
    const uint32_t owner_warp = NW == 1 ? warp : ((WR + 31) / 32 - 1) % NW;
    long long last[3] = {sm_last[owner_warp][0], sm_last[owner_warp][1], sm_last[owner_warp][2]};

    // ---- traceback (ConstructProfile, profile.cpp:727-775), lane 0 walks, the warp reverses
    uint32_t n = 0;
    long long total = 0;
    if (lane == 0) {
        int dir;
        if (last[0] >= last[1] && last[0] >= last[2]) { dir = 0; total = last[0]; }
        else if (last[1] > last[2]) { dir = 1; total = last[1]; }
        else { dir = 2; total = last[2]; }
        size_t i = WR, j = WC;
        while (i || j) {
            tmp_path[n++] = (unsigned char)dir;
            const unsigned char b = __ldcg(dirs + i * ld + j);
            if (dir == 0) { dir = b & 3; --i; --j; }
            else if (dir == 1) { dir = (b >> 2) & 3; --j; }
            else { dir = (b >> 4) & 3; --i; }
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
        r.swapped = (uint8_t)sw; r.variant = (uint8_t)var; r.pad[0] = r.pad[1] = 0;
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
    std::vector<uint32_t> order(n);
    unsigned long long path_off = 0, dirs_off = 0, scratch_off = 0, cells = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const famsa_dp_job& j = jobs[k];
        if (j.p1.width == 0 || j.p2.width == 0 || j.p1.card == 0 || j.p2.card == 0) {
            set_error("dp job " + std::to_string(k) + ": empty profile");
            return FAMSA_E_INVALID;
        }
        DpJobDev& d = dev[k];
        d.s1 = reinterpret_cast<const long long*>(j.p1.scores); d.c1 = j.p1.counters;
        d.s2 = reinterpret_cast<const long long*>(j.p2.scores); d.c2 = j.p2.counters;
        d.w1 = j.p1.width; d.card1 = j.p1.card; d.w2 = j.p2.width; d.card2 = j.p2.card;
        d.path_off = path_off; d.dirs_off = dirs_off; d.scratch_off = scratch_off;
        path_off += (unsigned long long)d.w1 + d.w2;
        dirs_off += ((unsigned long long)d.w1 + 1) * (d.w2 + 1);
        const unsigned long long wmax = std::max(d.w1, d.w2);
        scratch_off += ((sizeof(ColInfo) + sizeof(Cell)) * (wmax + 1) + d.w1 + d.w2 + 63) / 64 * 64;
        cells += (unsigned long long)d.w1 * d.w2;
    }
    // merges whose shorter side spans several 32-row stripes get a whole block (kDpTeamWarps warps pipelined
    // over the stripes); the rest run one warp per merge.  Both groups cost-descending.
    auto big = [&](uint32_t a) { return std::min(dev[a].w1, dev[a].w2) > kDpTeamMinWidth; };
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        if (big(a) != big(b)) return big(a);
        return (unsigned long long)dev[a].w1 * dev[a].w2 > (unsigned long long)dev[b].w1 * dev[b].w2;
    });
    uint32_t n_big = 0;
    while (n_big < n && big(order[n_big])) ++n_big;
    S.last_cells = cells;
    FB_TRY(S.d_jobs.reserve(sizeof(DpJobDev) * std::max(1u, n)));
    FB_TRY(S.d_order.reserve(sizeof(uint32_t) * std::max(1u, n)));
    FB_TRY(S.d_scratch.reserve(std::max<unsigned long long>(scratch_off, 64)));
    uint8_t* dirs = d_dirs;
    if (!dirs) {
        FB_TRY(S.d_dirs.reserve(std::max<unsigned long long>(dirs_off, 64)));
        dirs = S.d_dirs.as<uint8_t>();
    }
    FB_CUDA(cudaEventRecord(ctx->ev[0], st));
    FB_CUDA(cudaMemcpyAsync(S.d_jobs.p, dev.data(), sizeof(DpJobDev) * n, cudaMemcpyHostToDevice, st));
    FB_CUDA(cudaMemcpyAsync(S.d_order.p, order.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
    DpParams P{};
    P.jobs = S.d_jobs.as<DpJobDev>();
    P.order = S.d_order.as<uint32_t>();
    P.n_jobs = n;
    P.go = gaps[0]; P.ge = gaps[1]; P.to = gaps[2]; P.te = gaps[3];
    P.dirs = dirs;
    P.path = d_path;
    P.scratch = S.d_scratch.as<uint8_t>();
    P.results = d_results;
    FB_CUDA(cudaEventRecord(ctx->ev[1], st));
    // `order` is cost-descending with the team-kernel jobs first (see the sort above)
    static bool configured = false;
    constexpr size_t smem_big = (size_t)kDpTeamWarps * 2 * 30 * 32 * sizeof(int);
    constexpr size_t smem_small = (size_t)kDpWarps * 2 * 30 * 32 * sizeof(int);
    if (!configured) {
        FB_CUDA(cudaFuncSetAttribute(k_dp_align<kDpTeamWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big));
        FB_CUDA(cudaFuncSetAttribute(k_dp_align<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_small));
        configured = true;
    }
    if (n_big) {
        k_dp_align<kDpTeamWarps><<<n_big, kDpTeamWarps * 32, smem_big, st>>>(P);
        FB_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    if (n > n_big) {
        DpParams Q = P;
        Q.order = P.order + n_big;
        Q.n_jobs = n - n_big;
        k_dp_align<1><<<(Q.n_jobs + kDpWarps - 1) / kDpWarps, kDpWarps * 32, smem_small, st>>>(Q);
        FB_CUDA(cudaGetLastError());
        ctx->launches++;
    }
    FB_CUDA(cudaEventRecord(ctx->ev[2], st));
    FB_CUDA(cudaEventRecord(ctx->ev[3], st));
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
    S.h_stage.resize(bytes);
    FB_TRY(S.d_tables.reserve(std::max<unsigned long long>(bytes, 64)));
    std::vector<famsa_dp_job> dj(jobs, jobs + n);
    unsigned long long at = 0;
    uint8_t* hb = S.h_stage.data();
    uint8_t* db = S.d_tables.as<uint8_t>();
    auto put = [&](const void* src, size_t sz) { memcpy(hb + at, src, sz); void* d = db + at; at += sz; return d; };
    for (uint32_t k = 0; k < n; ++k) {      // all int64 tables first keeps 8-byte alignment trivially: sizes are multiples of 128
        famsa_dp_job& j = dj[k];
        j.p1.scores = static_cast<const int64_t*>(put(jobs[k].p1.scores, ((size_t)jobs[k].p1.width + 1) * 32 * 8));
        j.p2.scores = static_cast<const int64_t*>(put(jobs[k].p2.scores, ((size_t)jobs[k].p2.width + 1) * 32 * 8));
        j.p1.counters = static_cast<const int32_t*>(put(jobs[k].p1.counters, ((size_t)jobs[k].p1.width + 1) * 32 * 4));
        j.p2.counters = static_cast<const int32_t*>(put(jobs[k].p2.counters, ((size_t)jobs[k].p2.width + 1) * 32 * 4));
    }
    FB_TRY(S.d_results.reserve(sizeof(famsa_dp_result) * std::max(1u, n)));
    FB_TRY(S.d_path.reserve(std::max<unsigned long long>(path_total, 64)));
    uint8_t* d_dirs = nullptr;
    if (dirs_buf) {
        FB_TRY(S.d_dirs.reserve(std::max<unsigned long long>(dirs_total, 64)));
        d_dirs = S.d_dirs.as<uint8_t>();
    }
    if (bytes) FB_CUDA(cudaMemcpyAsync(db, hb, bytes, cudaMemcpyHostToDevice, st));
    FB_TRY(dp_run_device(ctx, dj.data(), n, gaps, S.d_results.as<famsa_dp_result>(), S.d_path.as<uint8_t>(), d_dirs, st));
    if (n) FB_CUDA(cudaMemcpyAsync(results, S.d_results.p, sizeof(famsa_dp_result) * n, cudaMemcpyDeviceToHost, st));
    if (path_total) FB_CUDA(cudaMemcpyAsync(path_buf, S.d_path.p, path_total, cudaMemcpyDeviceToHost, st));
    if (dirs_buf && dirs_total) FB_CUDA(cudaMemcpyAsync(dirs_buf, d_dirs, dirs_total, cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    return FAMSA_OK;
}

} // namespace fb
