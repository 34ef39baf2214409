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
    std::printf("host mirror ok\n");
    return 0;
}
