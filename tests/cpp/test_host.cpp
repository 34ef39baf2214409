// Exercises the C++ host mirror (famsa_b200/host/famsa_b200_host.hpp) the way the reference's tree
// builders use CLCSBP + calculateDistance*: results are compared with the oracle (liboracle.so).
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../famsa_b200/host/famsa_b200_host.hpp"

extern "C" {
void lcs_oracle_rows(const int8_t*, const uint64_t*, const uint32_t*, const uint32_t*, uint32_t, const uint32_t*, uint32_t, uint32_t*);
double lcs_oracle_transform_f64(int, uint32_t, uint32_t, uint32_t);
float lcs_oracle_transform_f32(int, uint32_t, uint32_t, uint32_t);
typedef struct { const int64_t* scores; const int32_t* counters; uint32_t width, card; } dp_oracle_profile;
int dp_oracle_align(const dp_oracle_profile*, const dp_oracle_profile*, const int64_t*, const int64_t*, uint8_t*, uint8_t*, uint32_t*,
                    int64_t*, int64_t*, int*, int*);
int dp_oracle_construct(const dp_oracle_profile*, const dp_oracle_profile*, const uint8_t*, uint32_t, const int64_t*, int64_t*, int32_t*,
                        uint32_t*, uint32_t*, uint32_t*, uint32_t*);
}

// host tables of a one-sequence profile (CProfile::CalculateCountersScores, profile.cpp:101-231)
struct HostProfile {
    std::vector<int64_t> s; std::vector<int32_t> c; uint32_t w, card;
    dp_oracle_profile view() const { return {s.data(), c.data(), w, card}; }
};
static HostProfile leaf_tables(const int8_t* seq, uint32_t len, const int64_t* sm, const int64_t* g)
{
    HostProfile p; p.w = len; p.card = 1; p.s.assign((size_t)(len + 1) * 32, 0); p.c.assign((size_t)(len + 1) * 32, 0);
    for (uint32_t col = 0; col <= len; ++col) {
        int64_t* s = &p.s[(size_t)col * 32];
        s[25] = g[0]; s[26] = g[1]; s[28] = g[2]; s[27] = g[3];
        if (col) { p.c[(size_t)col * 32 + seq[col - 1]] = 1; for (int k = 0; k < 24; ++k) s[k] = sm[seq[col - 1] * 24 + k]; }
    }
    return p;
}

#define REQUIRE(c) do { if (!(c)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main()
{
    using namespace famsa_b200;
    std::mt19937 rng(7);
    const uint32_t n = 150;
    std::vector<std::vector<int8_t>> seqs(n);
    for (auto& s : seqs) { s.resize(40 + rng() % 200); for (auto& c : s) c = (int8_t)(rng() % 21 == 20 ? 22 : rng() % 20); }
    std::sort(seqs.begin(), seqs.end(), [](auto& a, auto& b) { return a.size() > b.size(); });
    std::vector<SequenceView> views(n);
    std::vector<int8_t> flat; std::vector<uint64_t> off(n); std::vector<uint32_t> lens(n);
    for (uint32_t i = 0; i < n; ++i) { off[i] = flat.size(); lens[i] = (uint32_t)seqs[i].size(); flat.insert(flat.end(), seqs[i].begin(), seqs[i].end()); }
    for (uint32_t i = 0; i < n; ++i) views[i] = {flat.data() + off[i], lens[i]};

    Context ctx(0);
    CLCSBP lcsbp(ctx, views.data(), n);

    // calculateDistanceVector (SLINK/UPGMA/NJ row shape) with the default float transform
    Transform<float, Distance::indel075_div_lcs> tr;
    for (uint32_t row : {1u, 77u, 149u}) {
        std::vector<float> d(row);
        lcsbp.calculateDistanceVector(tr, row, row, d.data());
        std::vector<uint32_t> want(row);
        lcs_oracle_rows(flat.data(), off.data(), lens.data(), &row, 1, nullptr, row, want.data());
        for (uint32_t k = 0; k < row; ++k) REQUIRE(d[k] == lcs_oracle_transform_f32(0, want[k], lens[row], lens[k]));
    }
    // calculateDistanceRange (Prim / medoid shape) with the double transform
    Transform<double, Distance::indel075_div_lcs> trd;
    std::vector<uint32_t> ids = {5, 3, 140, 9, 9, 0};
    std::vector<double> dd(ids.size());
    lcsbp.calculateDistanceRange(trd, 42u, ids.begin(), ids.end(), dd.data());
    uint32_t ref = 42; std::vector<uint32_t> want(ids.size());
    lcs_oracle_rows(flat.data(), off.data(), lens.data(), &ref, 1, ids.data(), (uint32_t)ids.size(), want.data());
    for (size_t k = 0; k < ids.size(); ++k) REQUIRE(dd[k] == lcs_oracle_transform_f64(0, want[k], lens[42], lens[ids[k]]));
    // calculateDistanceMatrix (UPGMA partial trees) with pairwise identity
    Transform<float, Distance::pairwise_identity> pid;
    std::vector<float> tri((size_t)n * (n - 1) / 2);
    lcsbp.calculateDistanceMatrix(pid, n, tri.data());
    for (uint32_t i : {1u, 60u, 149u}) {
        std::vector<uint32_t> w(i);
        lcs_oracle_rows(flat.data(), off.data(), lens.data(), &i, 1, nullptr, i, w.data());
        for (uint32_t j = 0; j < i; ++j) REQUIRE(tri[(size_t)i * (i - 1) / 2 + j] == lcs_oracle_transform_f32(2, w[j], lens[i], lens[j]));
    }
    // raw seam
    uint32_t dist[8]; uint32_t eight[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    lcsbp.GetLCSBP(100, eight, 8, dist);
    uint32_t r100 = 100; uint32_t w8[8];
    lcs_oracle_rows(flat.data(), off.data(), lens.data(), &r100, 1, eight, 8, w8);
    for (int k = 0; k < 8; ++k) REQUIRE(dist[k] == w8[k]);
    // error convention: std::runtime_error
    bool threw = false;
    try { uint32_t bad = n; lcsbp.GetLCSBP(bad, eight, 8, dist); } catch (const std::runtime_error&) { threw = true; }
    REQUIRE(threw);
    // ---- the default guide tree's MST on the device: n-1 edges, every sequence visited exactly once
    {
        PrimEdges e = PrimMST(ctx, Distance::indel075_div_lcs);
        REQUIRE(e.from.size() == n - 1);
        std::vector<int> seen(n, 0);
        for (int32_t o : e.order) { REQUIRE(o >= 0 && (uint32_t)o < n); ++seen[o]; }
        for (uint32_t i = 0; i < n; ++i) REQUIRE(seen[i] == 1);
        for (size_t k = 0; k < e.from.size(); ++k) {
            REQUIRE(e.from[k] < e.to[k]);
            uint32_t a = (uint32_t)e.from[k], b = (uint32_t)e.to[k], l;
            // distance of the pair with the later-visited... both orientations agree on sequences without the carry quirk
            lcs_oracle_rows(flat.data(), off.data(), lens.data(), &a, 1, &b, 1, &l);
            REQUIRE(e.dist[k] == lcs_oracle_transform_f64(0, l, lens[a], lens[b]));
        }
    }
    // ---- resident profiles: two leaf merges, then the merge of their results, against the restatement
    {
        std::vector<int64_t> sm(24 * 24);
        for (int a = 0; a < 24; ++a) for (int b = 0; b <= a; ++b) sm[a * 24 + b] = sm[b * 24 + a] = (a == b ? 5000 + 100 * a : (int64_t)((a * 7 + b * 13) % 9) * 500 - 2500);
        const int64_t g[4] = {-14850, -1250, -660, -660};
        ResidentProfiles rp(ctx, sm.data(), lens);
        std::vector<uint32_t> ids;
        std::vector<famsa_prof_merge> lvl = {{ResidentProfiles::Leaf(3), ResidentProfiles::Leaf(4)}, {ResidentProfiles::Leaf(120), ResidentProfiles::Leaf(7)}};
        auto r1 = rp.MergeLevel(lvl, g, ids);
        std::vector<HostProfile> merged;
        const uint32_t leaves[2][2] = {{3, 4}, {120, 7}};
        for (int k = 0; k < 2; ++k) {
            HostProfile a = leaf_tables(flat.data() + off[leaves[k][0]], lens[leaves[k][0]], sm.data(), g);
            HostProfile b = leaf_tables(flat.data() + off[leaves[k][1]], lens[leaves[k][1]], sm.data(), g);
            dp_oracle_profile pa = a.view(), pb = b.view();
            std::vector<uint8_t> dirs((size_t)(a.w + 1) * (b.w + 1)), path(a.w + b.w + 1);
            uint32_t plen; int64_t last[3], total; int sw, var;
            dp_oracle_align(&pa, &pb, g, nullptr, dirs.data(), path.data(), &plen, last, &total, &sw, &var);
            REQUIRE(plen == r1[k].path.size() && total == r1[k].total_score && (sw != 0) == r1[k].swapped);
            REQUIRE(std::equal(path.begin(), path.begin() + plen, r1[k].path.begin()));
            HostProfile m; m.w = plen; m.card = 2; m.s.resize((size_t)(plen + 1) * 32); m.c.resize((size_t)(plen + 1) * 32);
            std::vector<uint32_t> g1(2 * plen + 2), g2(2 * plen + 2); uint32_t n1, n2;
            dp_oracle_profile R = sw ? pb : pa, C = sw ? pa : pb;
            REQUIRE(dp_oracle_construct(&R, &C, path.data(), plen, g, m.s.data(), m.c.data(), g1.data(), &n1, g2.data(), &n2) == 0);
            auto runs = GapRuns(r1[k].path, 1);
            REQUIRE(runs.size() == n1);
            for (uint32_t q = 0; q < n1; ++q) REQUIRE(runs[q].first == g1[2 * q] && runs[q].second == g1[2 * q + 1]);
            std::vector<int64_t> ds; std::vector<int32_t> dc; uint32_t card;
            rp.Download(ids[k], ds, dc, card);
            REQUIRE(card == 2 && ds == m.s && dc == m.c);
            merged.push_back(std::move(m));
        }
        std::vector<uint32_t> top;
        auto r2 = rp.MergeLevel({{ids[0], ids[1]}}, g, top);
        dp_oracle_profile pa = merged[0].view(), pb = merged[1].view();
        std::vector<uint8_t> dirs((size_t)(pa.width + 1) * (pb.width + 1)), path(pa.width + pb.width + 1);
        uint32_t plen; int64_t last[3], total; int sw, var;
        dp_oracle_align(&pa, &pb, g, nullptr, dirs.data(), path.data(), &plen, last, &total, &sw, &var);
        REQUIRE(var == 2 && plen == r2[0].path.size() && total == r2[0].total_score);
        REQUIRE(std::equal(path.begin(), path.begin() + plen, r2[0].path.begin()));
        REQUIRE(rp.Width(top[0]) == plen);
        rp.Drop(top);
        uint64_t live = 1, bytes = 1;
        check(famsa_prof_stats(ctx.get(), &live, &bytes));
        REQUIRE(live == 0 && bytes == 0);
    }
    // ---- the whole merge loop in one call (AlignTree) against the same tree run level by level (ResidentProfiles)
    {
        std::vector<int64_t> sm(24 * 24);
        for (int a = 0; a < 24; ++a) for (int b = 0; b <= a; ++b) sm[a * 24 + b] = sm[b * 24 + a] = (a == b ? 5000 + 100 * a : (int64_t)((a * 7 + b * 13) % 9) * 500 - 2500);
        const int64_t g[4] = {-14850, -1250, -660, -660};
        // the UPGMA tree of the set, built on the device, as the guide tree: n leaf entries + n - 1 merges
        auto up = UPGMATree(ctx, Distance::indel075_div_lcs, false);
        REQUIRE(up.size() == n - 1);
        std::vector<std::pair<int, int>> tree(n, std::make_pair(-1, -1));
        tree.insert(tree.end(), up.begin(), up.end());
        std::vector<int> used(2 * n - 1, 0);
        for (uint32_t k = 0; k + 1 < n; ++k) {
            REQUIRE(up[k].first >= 0 && up[k].second >= 0 && (uint32_t)up[k].first < n + k && (uint32_t)up[k].second < n + k);
            ++used[up[k].first]; ++used[up[k].second];
        }
        for (uint32_t v = 0; v + 1 < 2 * n - 1; ++v) REQUIRE(used[v] == 1);                // every node but the root is merged exactly once
        ResidentProfiles rp(ctx, sm.data(), lens);
        TreeAlignment ta = AlignTree(ctx, tree, n, g);
        REQUIRE(ta.merges.size() == n - 1 && ta.stats.cells > 0);
        std::vector<uint32_t> node(2 * n - 1);
        for (uint32_t i = 0; i < n; ++i) node[i] = ResidentProfiles::Leaf(i);
        for (auto& level : ReadyLevels(tree, n)) {
            std::vector<famsa_prof_merge> lvl;
            for (uint32_t k : level) lvl.push_back({node[tree[n + k].first], node[tree[n + k].second]});
            std::vector<uint32_t> ids;
            auto r = rp.MergeLevel(lvl, g, ids);
            for (size_t q = 0; q < level.size(); ++q) {
                const AlignResult& a = ta.merges[level[q]];
                REQUIRE(a.total_score == r[q].total_score && a.path == r[q].path && a.swapped == r[q].swapped);
                node[n + level[q]] = ids[q];
            }
        }
        uint32_t w = 0, card = 0;
        check(famsa_prof_get(ctx.get(), ta.root_id, &w, &card, nullptr, nullptr));
        REQUIRE(card == n && w == ta.merges.back().path.size());
        rp.Drop({node[2 * n - 2], ta.root_id});
        uint64_t live = 1, bytes = 1;
        check(famsa_prof_stats(ctx.get(), &live, &bytes));
        REQUIRE(live == 0 && bytes == 0);
    }
    std::printf("host mirror ok\n");
    return 0;
}
