/*
 * oracle/dp_oracle.c -- CPU restatement of the reference's profile-alignment DP (HP-2).
 *
 * TEST INFRASTRUCTURE ONLY (same rules as lcs_oracle.c).
 *
 * Parity: PINNED.  tests/test_oracle_dp.py checks this restatement against
 *   - the reference's golden test/adeno_fiber/upgma.pp.fasta (one ProfProf merge of the two
 *     halves upgma.no_refine.part{1,2}.fasta; fixture tests/golden/adeno_pp.npz),
 *   - the reference itself (oracle/_ref: CProfile::Align + ConstructProfile, sequential and
 *     2-thread variants) on every merge of guide trees over adeno_fiber / hemopexin subsets and
 *     random families: traceback path and total score must be identical.
 *
 * What each function follows (paths relative to /root/reference/src/core):
 *   dp_gap_start / dp_gap_cont    profile.cpp:1223-1278 / 1281-1315 (DP_SolveGapsProblemWhen*)
 *   dp_oracle_align (dispatch)    profile.cpp:244-305 (CProfile::Align)
 *   cell loops, ProfProf          profile_par.cpp:441-903 (== profile_seq.cpp:495-892 unbanded)
 *              SeqProf            profile_par.cpp:26-438  (== profile_seq.cpp:165-491 unbanded)
 *              SeqSeq             profile_seq.cpp:24-162
 *   traceback                     profile.cpp:727-782 (first part of ConstructProfile)
 *
 * All arithmetic is int64 (score_t, defs.h:36-40), NEG = -(1<<62) takes part in additions
 * unsaturated (defs.h:57).  Direction byte = dirD | dirH<<2 | dirV<<4, D=0 H=1 V=2 (profile.h:33,93-142).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NSYM 32
#define GO 25   /* GAP_OPEN */
#define GE 26   /* GAP_EXT */
#define TE 27   /* GAP_TERM_EXT */
#define TO 28   /* GAP_TERM_OPEN */
#define NEG (-(1ll << 62))
enum { DIR_D = 0, DIR_H = 1, DIR_V = 2 };

typedef struct {
    const int64_t *scores;    /* (width+1) x 32, column-major (CProfileValues, profile.h:153) */
    const int32_t *counters;  /* (width+1) x 32 */
    uint32_t width, card;
} dp_oracle_profile;

typedef struct { int64_t s_o, s_e, s_to, s_te, k_e, k_te; } gapcorr;
typedef struct { int64_t D, H, V; } cell;

static gapcorr solve_gaps(const dp_oracle_profile *p, uint32_t c)
{
    const int32_t *cc = p->counters + (size_t)c * NSYM;
    const int64_t n = p->card;
    gapcorr g = {0, 0, 0, 0, 0, 0};
    if (c >= p->width) {                              /* profile.cpp:1234-1249 */
        g.s_te = cc[TO] + cc[TE];
        g.s_to = n - g.s_te;
        g.k_te = n;                                   /* :1291-1296 */
        g.k_e = 0;
    } else {
        const int32_t *cn = cc + NSYM;
        g.s_to = cn[TO];                              /* :1255-1273 */
        g.s_te = cc[TO] + cc[TE];
        g.s_e = cc[GO] + cc[GE];
        g.s_o = n - g.s_e - g.s_to - g.s_te;
        g.k_te = (int64_t)cn[TO] + cc[TO] + cc[TE];   /* :1301-1310 */
        g.k_e = n - g.k_te;
    }
    return g;
}

static size_t count_nonzero(const dp_oracle_profile *p)
{
    size_t nz = 0;
    for (size_t i = 0; i < ((size_t)p->width + 1) * NSYM; ++i) nz += p->counters[i] != 0;
    return nz;
}

/* the residue of column i of a one-sequence profile */
static int seq_symbol(const dp_oracle_profile *p, uint32_t i)
{
    const int32_t *cc = p->counters + (size_t)i * NSYM;
    for (int k = 0; k < 24; ++k)
        if (cc[k]) return k;
    return 22;
}

static inline uint8_t pick3(int64_t a, int64_t b, int64_t c, int da, int db, int dc, int64_t *out)
{   /* a if a>b && a>c; else b if b>c; else c   (strict, fixed priority) */
    if (a > b && a > c) { *out = a; return (uint8_t)da; }
    if (b > c) { *out = b; return (uint8_t)db; }
    *out = c; return (uint8_t)dc;
}

/* variant: 0 SeqSeq, 1 SeqProf, 2 ProfProf.  R = rows, C = columns. */
static void dp_fill(int variant, const dp_oracle_profile *R, const dp_oracle_profile *C,
                    const int64_t gaps[4], const int64_t *score_matrix /* 24x24 or NULL */,
                    uint8_t *dirs, int64_t last[3])
{
    const uint32_t WR = R->width, WC = C->width;
    const int64_t go = gaps[0], ge = gaps[1], to = gaps[2], te = gaps[3];
    const int64_t nR = R->card, nC = C->card;
    const size_t ld = (size_t)WC + 1;
    cell *prev = (cell *)malloc(sizeof(cell) * ld), *cur = (cell *)malloc(sizeof(cell) * ld);
    gapcorr *g2 = (gapcorr *)calloc(ld, sizeof(gapcorr));
    int64_t *chg2 = (int64_t *)calloc(ld, sizeof(int64_t));
    int *seqC = (int *)calloc(ld, sizeof(int));
    memset(dirs, 0, ((size_t)WR + 1) * ld);

    for (uint32_t j = 1; j <= WC; ++j) {
        g2[j] = solve_gaps(C, j);
        const int32_t *cc = C->counters + (size_t)j * NSYM;
        chg2[j] = cc[GO] * (ge - go) + cc[TO] * (te - to);
        seqC[j] = seq_symbol(C, j);
    }
#define SC(j, k) (C->scores[(size_t)(j) * NSYM + (k)])
#define SR(i, k) (R->scores[(size_t)(i) * NSYM + (k)])

    /* row 0 */
    prev[0].D = 0; prev[0].H = NEG; prev[0].V = NEG;
    for (uint32_t j = 1; j <= WC; ++j) {
        prev[j].D = NEG; prev[j].V = NEG;
        if (variant == 0) prev[j].H = j == 1 ? to : (prev[j - 1].H > prev[j - 1].D ? prev[j - 1].H : prev[j - 1].D) + te;
        else if (variant == 1) prev[j].H = j == 1 ? prev[0].D + SC(1, TO) : prev[j - 1].H + SC(j, TE);
        else prev[j].H = j == 1 ? prev[0].D + SC(1, TO) * nR : prev[j - 1].H + SC(j, TE) * nR;
        dirs[j] = DIR_H | DIR_H << 2 | DIR_H << 4;
    }
    prev[WC].H = NEG;

    for (uint32_t i = 1; i <= WR; ++i) {
        const int32_t *rc = R->counters + (size_t)i * NSYM;
        uint8_t *drow = dirs + (size_t)i * ld;
        const int last_row = i == WR;
        /* column 0 */
        cur[0].D = NEG; cur[0].H = NEG;
        drow[0] = DIR_V | DIR_V << 2 | DIR_V << 4;
        if (!last_row) {
            int64_t m = prev[0].D > prev[0].V ? prev[0].D : prev[0].V;
            int64_t cost;
            if (variant == 2) cost = (i == 1 ? SR(i, TO) : SR(i, TE)) * nC;
            else if (variant == 1) cost = (i == 1 ? to : te) * nC;
            else cost = i == 1 ? to : te;
            cur[0].V = m + cost;
        } else
            cur[0].V = NEG;

        /* per-row constants */
        const int symR = seq_symbol(R, i);
        gapcorr g1 = {0, 0, 0, 0, 0, 0};
        int64_t g1o = 0, g1t = 0, nongap1 = 0;
        int nzk[30], nzn = 0;
        int64_t nzc[30];
        if (variant == 2) {
            g1 = solve_gaps(R, i);
            g1o = rc[GO]; g1t = rc[TO];
            for (int k = 0; k < 30; ++k)
                if (rc[k]) { nzk[nzn] = k; nzc[nzn++] = rc[k]; if (k < 24) nongap1 += rc[k]; }
        }

        for (uint32_t j = 1; j <= WC; ++j) {
            const cell P = prev[j - 1], L = cur[j - 1], U = prev[j];
            const int three = i > 1 && j > 1;
            uint8_t dD, dH, dV;
            int64_t vD, vH, vV;
            if (variant == 0) {
                /* profile_seq.cpp:98-113: note the >= in the second test */
                int64_t s = score_matrix ? score_matrix[symR * 24 + seqC[j]] : SC(j, symR);
                if (P.D > P.H && P.D > P.V) { vD = P.D + s; dD = DIR_D; }
                else if (P.H >= P.V) { vD = P.H + s; dD = DIR_H; }
                else { vD = P.V + s; dD = DIR_V; }
                int64_t tD = L.D + (!last_row ? go : to), tH = L.H + (!last_row ? ge : te);
                if (tD > tH) { vH = tD; dH = DIR_D; } else { vH = tH; dH = DIR_H; }
                tD = U.D + (j < WC ? go : to);
                int64_t tV = U.V + (j < WC ? ge : te);
                if (tD > tV) { vV = tD; dV = DIR_D; } else { vV = tV; dV = DIR_V; }
            } else if (variant == 1) {
                /* profile_par.cpp:255-421 */
                const int64_t t = SC(j, symR);
                dD = pick3(P.D, P.H, P.V + chg2[j], DIR_D, DIR_H, DIR_V, &vD);
                vD += t;
                const int64_t gcH = !last_row ? SC(j, GO) : SC(j, TO);
                int64_t tD = L.D + gcH, tH = L.H + (!last_row ? SC(j, GE) : SC(j, TE));
                if (three) dH = pick3(tD, L.V + gcH, tH, DIR_D, DIR_V, DIR_H, &vH);   /* D, then V > H, else H */
                else if (tD > tH) { vH = tD; dH = DIR_D; } else { vH = tH; dH = DIR_H; }
                const int64_t gcV = go * g2[j].s_o + ge * g2[j].s_e + to * g2[j].s_to + te * g2[j].s_te;
                tD = U.D + gcV;
                int64_t tV = U.V + ge * g2[j].k_e + te * g2[j].k_te;
                if (three) dV = pick3(tD, U.H + gcV, tV, DIR_D, DIR_H, DIR_V, &vV);
                else if (tD > tV) { vV = tD; dV = DIR_D; } else { vV = tV; dV = DIR_V; }
            } else {
                /* profile_par.cpp:679-886 */
                int64_t t = 0;
                for (int q = 0; q < nzn; ++q) t += nzc[q] * SC(j, nzk[q]);
                int64_t tD = P.D + t;
                int64_t tH = P.H + t;
                if (g1o || g1t) tH += g1o * (SC(j, GE) - SC(j, GO)) + g1t * (SC(j, TE) - SC(j, TO));
                int64_t tV = P.V + t + chg2[j] * nongap1;
                dD = pick3(tD, tH, tV, DIR_D, DIR_H, DIR_V, &vD);
                const int64_t gcH = SC(j, GO) * g1.s_o + SC(j, GE) * g1.s_e + SC(j, TO) * g1.s_to + SC(j, TE) * g1.s_te;
                tD = L.D + gcH;
                tH = L.H + SC(j, GE) * g1.k_e + SC(j, TE) * g1.k_te;
                if (three) dH = pick3(tD, L.V + gcH, tH, DIR_D, DIR_V, DIR_H, &vH);
                else if (tD > tH) { vH = tD; dH = DIR_D; } else { vH = tH; dH = DIR_H; }
                const int64_t gcV = SR(i, GO) * g2[j].s_o + SR(i, GE) * g2[j].s_e + SR(i, TO) * g2[j].s_to + SR(i, TE) * g2[j].s_te;
                tD = U.D + gcV;
                tV = U.V + SR(i, GE) * g2[j].k_e + SR(i, TE) * g2[j].k_te;
                if (three) dV = pick3(tD, U.H + gcV, tV, DIR_D, DIR_H, DIR_V, &vV);
                else if (tD > tV) { vV = tD; dV = DIR_D; } else { vV = tV; dV = DIR_V; }
            }
            cur[j].D = vD; cur[j].H = vH; cur[j].V = vV;
            drow[j] = (uint8_t)(dD | dH << 2 | dV << 4);
        }
        cell *t = prev; prev = cur; cur = t;
    }
    last[0] = prev[WC].D; last[1] = prev[WC].H; last[2] = prev[WC].V;
    free(prev); free(cur); free(g2); free(chg2); free(seqC);
#undef SC
#undef SR
}

/* profile.cpp:727-782.  path[0..*path_len) = moves in forward order (0 D: row+col, 1 H: column
 * only / gap in the row profile, 2 V: row only). */
static void dp_traceback(const uint8_t *dirs, uint32_t WR, uint32_t WC, const int64_t last[3],
                         uint8_t *path, uint32_t *path_len, int64_t *total)
{
    int dir;
    if (last[0] >= last[1] && last[0] >= last[2]) { dir = DIR_D; *total = last[0]; }
    else if (last[1] > last[2]) { dir = DIR_H; *total = last[1]; }
    else { dir = DIR_V; *total = last[2]; }
    size_t i = WR, j = WC, n = 0;
    const size_t ld = (size_t)WC + 1;
    while (i || j) {
        path[n++] = (uint8_t)dir;
        const uint8_t b = dirs[i * ld + j];
        if (dir == DIR_D) { dir = b & 3; --i; --j; }
        else if (dir == DIR_H) { dir = (b >> 2) & 3; --j; }
        else { dir = (b >> 4) & 3; --i; }
    }
    for (size_t a = 0, b = n ? n - 1 : 0; a < b; ++a, --b) { uint8_t t = path[a]; path[a] = path[b]; path[b] = t; }
    *path_len = (uint32_t)n;
}

/* CProfile::Align: variant + orientation, fill, traceback.
 * dirs: caller buffer of (max(W1,W2)+1)^2... precisely (WR+1)*(WC+1) for the chosen orientation,
 * so allocate (W1+1)*(W2+1) (the product does not depend on the orientation).
 * force: -1 = reference dispatch; otherwise bit0 = swapped. */
int dp_oracle_align(const dp_oracle_profile *p1, const dp_oracle_profile *p2, const int64_t gaps[4],
                    const int64_t *score_matrix, uint8_t *dirs, uint8_t *path, uint32_t *path_len,
                    int64_t last[3], int64_t *total, int *swapped, int *variant)
{
    const dp_oracle_profile *R = p1, *C = p2;
    int var, sw = 0;
    if (p1->card == 1 && p2->card == 1) var = 0;
    else if (p1->card == 1) var = 1;
    else if (p2->card == 1) { var = 1; sw = 1; }
    else {
        var = 2;
        if (!(count_nonzero(p1) * (size_t)p2->width < count_nonzero(p2) * (size_t)p1->width)) sw = 1;
    }
    if (sw) { R = p2; C = p1; }
    dp_fill(var, R, C, gaps, score_matrix, dirs, last);
    dp_traceback(dirs, R->width, C->width, last, path, path_len, total);
    *swapped = sw;
    *variant = var;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Merge part of ConstructProfile: given the traceback path, build the merged profile's
 * scores/counters and the gap-run lists that FinalizeGaps applies to the members of either child.
 * Follows profile.cpp:784-1002 (the walk over the path), :1005-1050 (InsertGaps), :1107-1111
 * (InsertColumn), :1114-1143 / :1146-1220 (SolveGapsProblemWhenContinuing / WhenStarting), written the
 * way the reference runs it: one sequential walk that carries the pending "open becomes ext" transfers
 * and mutates (a private copy of) the children's tables.  R = rows of the DP matrix (ConstructProfile's
 * profile1), C = columns (profile2); path = direction_t bytes in forward order (path[1..width]).
 *
 * Pinned by tests/test_oracle_dp.py against the reference's own ConstructProfile (oracle/_ref) on every
 * merge of the adeno_fiber / hemopexin guide trees and random families.
 * out_scores/out_counters: (plen+1) x 32.  gaps1/gaps2: (first column, run length) pairs, at most plen each.
 */
typedef struct { int32_t to_transfer, term_to_transfer, o_left, e_left, to_left, te_left; } walk_side;

static void apply_transfer(int64_t *s, int32_t *c, uint32_t col, walk_side *w, const int64_t g[4])
{
    if (!w->to_transfer && !w->term_to_transfer) return;
    c[(size_t)col * NSYM + GE] += w->to_transfer;       c[(size_t)col * NSYM + GO] -= w->to_transfer;
    c[(size_t)col * NSYM + TE] += w->term_to_transfer;  c[(size_t)col * NSYM + TO] -= w->term_to_transfer;
    const int64_t cost = w->to_transfer * (g[1] - g[0]) + w->term_to_transfer * (g[3] - g[2]);
    for (int k = 0; k < 24; ++k) s[(size_t)col * NSYM + k] += cost;
    w->to_transfer = w->term_to_transfer = 0;
}

int dp_oracle_construct(const dp_oracle_profile *R, const dp_oracle_profile *Cc, const uint8_t *path, uint32_t plen,
                        const int64_t gaps[4], int64_t *out_scores, int32_t *out_counters,
                        uint32_t *gaps1, uint32_t *n_gaps1, uint32_t *gaps2, uint32_t *n_gaps2)
{
    const size_t n1 = ((size_t)R->width + 1) * NSYM, n2 = ((size_t)Cc->width + 1) * NSYM;
    int64_t *s1 = malloc(n1 * 8), *s2 = malloc(n2 * 8);
    int32_t *c1 = malloc(n1 * 4), *c2 = malloc(n2 * 4);
    if (!s1 || !s2 || !c1 || !c2) return -1;
    memcpy(s1, R->scores, n1 * 8); memcpy(c1, R->counters, n1 * 4);
    memcpy(s2, Cc->scores, n2 * 8); memcpy(c2, Cc->counters, n2 * 4);
    memset(out_scores, 0, ((size_t)plen + 1) * NSYM * 8);
    memset(out_counters, 0, ((size_t)plen + 1) * NSYM * 4);
    walk_side w1 = {0}, w2 = {0};
    uint32_t i = 0, j = 0, run = 0;
    *n_gaps1 = *n_gaps2 = 0;
    int prev = DIR_D;
    for (uint32_t k = 1; k <= plen; ++k) {
        const int dir = path[k - 1];
        const int next = k < plen ? path[k] : -1;          /* the reference appends a V that never continues a run */
        int64_t *os = out_scores + (size_t)k * NSYM;
        int32_t *oc = out_counters + (size_t)k * NSYM;
        if (dir == DIR_D) {
            ++i; ++j;
            apply_transfer(s1, c1, i, &w1, gaps);
            apply_transfer(s2, c2, j, &w2, gaps);
            w1.o_left = w1.e_left = w1.to_left = w1.te_left = 0;
            w2.o_left = w2.e_left = w2.to_left = w2.te_left = 0;
            for (int r = 0; r < NSYM; ++r) {
                oc[r] += c1[(size_t)i * NSYM + r] + c2[(size_t)j * NSYM + r];
                os[r] += s1[(size_t)i * NSYM + r] + s2[(size_t)j * NSYM + r];
            }
        } else {
            /* gap column inserted into G (rows side for H, columns side for V); the other child supplies a column */
            const int isH = dir == DIR_H;
            const dp_oracle_profile *G = isH ? R : Cc;
            int32_t *cg = isH ? c1 : c2;
            walk_side *wg = isH ? &w1 : &w2;
            const uint32_t src = isH ? i : j, W = G->width;
            const int32_t size = (int32_t)G->card;
            int32_t o = 0, e = 0, to = 0, te = 0;
            if (prev == dir) {                              /* SolveGapsProblemWhenContinuing */
                if (src == W || src == 0) te += size;
                else {
                    te += wg->to_left; te += wg->te_left;
                    e = wg->o_left; e += wg->e_left;
                    o = size - e - te;
                }
            } else {                                        /* SolveGapsProblemWhenStarting */
                if (src == 0) {
                    to += size;
                    wg->term_to_transfer = cg[(size_t)(src + 1) * NSYM + TO];
                } else if (src >= W) {
                    const int32_t cnt = cg[(size_t)src * NSYM + TO] + cg[(size_t)src * NSYM + TE];
                    te = cnt; to += size - cnt;
                } else {
                    to += cg[(size_t)(src + 1) * NSYM + TO];
                    wg->term_to_transfer = to;
                    te += cg[(size_t)src * NSYM + TO]; te += cg[(size_t)src * NSYM + TE];
                    e = cg[(size_t)src * NSYM + GO]; e += cg[(size_t)src * NSYM + GE];
                    o = cg[(size_t)(src + 1) * NSYM + GO];
                    wg->to_transfer += o;
                    o = size - e - to - te;
                }
            }
            wg->o_left = o; wg->e_left = e; wg->to_left = to; wg->te_left = te;
            /* InsertGaps: run-length list + the gap column's contribution */
            ++run;
            if (!(next == dir)) {
                uint32_t *lst = isH ? gaps1 : gaps2;
                uint32_t *cnt = isH ? n_gaps1 : n_gaps2;
                lst[2 * *cnt] = k + 1 - run; lst[2 * *cnt + 1] = run; ++*cnt;
                run = 0;
            }
            const int64_t cost = o * gaps[0] + e * gaps[1] + to * gaps[2] + te * gaps[3];
            oc[GO] += o; oc[GE] += e; oc[TO] += to; oc[TE] += te; oc[30] += size;
            for (int r = 0; r < 24; ++r) os[r] += cost;
            /* the other child's next column, after its own pending transfer */
            if (isH) {
                apply_transfer(s2, c2, j + 1, &w2, gaps);
                ++j;
                for (int r = 0; r < NSYM; ++r) { oc[r] += c2[(size_t)j * NSYM + r]; os[r] += s2[(size_t)j * NSYM + r]; }
            } else {
                apply_transfer(s1, c1, i + 1, &w1, gaps);
                ++i;
                for (int r = 0; r < NSYM; ++r) { oc[r] += c1[(size_t)i * NSYM + r]; os[r] += s1[(size_t)i * NSYM + r]; }
            }
        }
        prev = dir;
    }
    const int64_t tot = (int64_t)R->card + Cc->card;        /* profile.cpp:998-1001 */
    out_scores[GO] = gaps[0] * tot; out_scores[GE] = gaps[1] * tot;
    out_scores[TO] = gaps[2] * tot; out_scores[TE] = gaps[3] * tot;
    free(s1); free(s2); free(c1); free(c2);
    return (i == R->width && j == Cc->width) ? 0 : -2;
}
