// oracle/ref_harness.cpp -- thin C entry points over the UNMODIFIED reference sources.
//
// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  This translation unit is ours; it is linked
// against object files compiled straight from /root/reference/src (never copied into the repo)
// by oracle/Makefile, producing oracle/_ref/libfamsa_ref.so.  It is used
//   * by tests/ to pin oracle/lcs_oracle.c and oracle/dp_oracle.c against the real reference,
//   * by bench.py --impl reference / cpu_baseline (kind "reference") as the timed CPU arm.
// Nothing under famsa_b200/ may load it.
//
// Reference seams exercised (paths relative to /root/reference):
//   CSequence ctor / DataResize / ComputeBitMasks          src/core/sequence.cpp:22-201
//   CLCSBP::GetLCSBP (8-way CSequence* overload)           src/lcs/lcsbp.cpp:163-264
//   AbstractTreeGenerator::calculateDistanceVector         src/tree/AbstractTreeGenerator.hpp:131-182
//   Transform<T, Distance>                                 src/tree/AbstractTreeGenerator.hpp:28-82
//   CProfile(const CGappedSequence&, CParams*)             src/core/profile.cpp:45-51
//   CProfile(CProfile*, CProfile*, CParams*, ...) -> Align src/core/profile.cpp:69-75, 244-305
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#define private public      // harness TU only: read CProfile internals, nothing is modified
#define protected public
#include "core/defs.h"
#include "core/params.h"
#include "core/profile.h"
#include "core/sequence.h"
#include "lcs/lcsbp.h"
#include "tree/AbstractTreeGenerator.h"
#include "tree/AbstractTreeGenerator.hpp"
#undef private
#undef protected

namespace {

// Smallest concrete generator: we only want the batch-driver templates of the base class.
struct DriverOnly : public AbstractTreeGenerator {
    DriverOnly(instruction_set_t isa) : AbstractTreeGenerator(1, isa) {}
    void run(std::vector<CSequence*>&, tree_structure&) override {}
};

// "Transform" that hands back the raw LCS length instead of a distance.
struct RawLcs {
    uint32_t operator()(uint32_t lcs, uint32_t, uint32_t) { return lcs; }
};

instruction_set_t isa_from_int(int isa)
{
    switch (isa) {
    case 0: return instruction_set_t::none;      // scalar CLCSBP_Classic
    case 1: return instruction_set_t::avx;       // 2 pairs / call
    default: return instruction_set_t::avx2;     // 4 pairs / call (the metric's CPU baseline)
    }
}

struct SeqSet {
    std::vector<CSequence> seqs;       // padded to the set-wide max length with UNKNOWN (msa.cpp:271,297)
    std::vector<CSequence*> ptrs;
};

} // namespace

extern "C" {

// ---------------------------------------------------------------- sequence sets
// `letters[i]` is the raw residue string; encoding is the reference's own (sequence.cpp:53-79).
void* ref_seqset_create(const char* const* letters, uint32_t n)
{
    auto* s = new SeqSet();
    s->seqs.reserve(n);
    uint32_t max_len = 0;
    for (uint32_t i = 0; i < n; ++i) {
        s->seqs.emplace_back("s" + std::to_string(i), std::string(letters[i]), (int)i, nullptr);
        if (s->seqs.back().length > max_len) max_len = s->seqs.back().length;
    }
    for (auto& q : s->seqs)
        if (q.length) q.DataResize(max_len, UNKNOWN_SYMBOL);
    for (auto& q : s->seqs) s->ptrs.push_back(&q);
    return s;
}

void ref_seqset_destroy(void* h) { delete static_cast<SeqSet*>(h); }

uint32_t ref_seqset_len(void* h, uint32_t i) { return static_cast<SeqSet*>(h)->seqs[i].length; }

// copies the reference's residue codes (without padding) into out[0..len)
void ref_seqset_codes(void* h, uint32_t i, int8_t* out)
{
    auto& q = static_cast<SeqSet*>(h)->seqs[i];
    if (q.length) memcpy(out, q.data, q.length);
}

// ---------------------------------------------------------------- HP-1: LCS
// Row `ref` (supplies the bit masks) against sequences[0..n_cols) -- exactly what SLINK/UPGMA/NJ/
// DistanceCalculator do per row.  out[k] = LCS length.
void ref_lcs_row_prefix(void* h, uint32_t ref, uint32_t n_cols, uint32_t* out, int isa)
{
    auto* s = static_cast<SeqSet*>(h);
    DriverOnly drv(isa_from_int(isa));
    CLCSBP lcsbp(isa_from_int(isa));
    RawLcs raw;
    drv.calculateDistanceVector<CSequence*, uint32_t, RawLcs>(raw, s->ptrs[ref], s->ptrs.data(),
                                                              (int)n_cols, out, lcsbp);
}

// Row `ref` against an arbitrary id list (Prim / medoid shape, calculateDistanceRange).
void ref_lcs_row_ids(void* h, uint32_t ref, const int* ids, uint32_t n_ids, uint32_t* out, int isa)
{
    auto* s = static_cast<SeqSet*>(h);
    DriverOnly drv(isa_from_int(isa));
    CLCSBP lcsbp(isa_from_int(isa));
    RawLcs raw;
    s->ptrs[ref]->ComputeBitMasks();       // calculateDistanceRange expects the caller to do this
    drv.calculateDistanceRange<CSequence*, uint32_t, const int*, RawLcs>(
        raw, s->ptrs[ref], s->ptrs.data(), std::make_pair(ids, ids + n_ids), out, lcsbp);
    s->ptrs[ref]->ReleaseBitMasks();
}

// Multi-threaded triangle rows [row_begin,row_end) with the reference's default transform
// (float indel075_div_lcs, the UPGMA shape, UPGMA.cpp:75-109).  Returns wall seconds; *n_pairs is
// the number of LCS lengths produced.  If out_lcs != nullptr, raw LCS lengths are stored packed
// (row i at i(i-1)/2 - row_begin(row_begin-1)/2) instead of being transformed.
double ref_lcs_triangle_mt(void* h, uint32_t row_begin, uint32_t row_end, int n_threads, int isa,
                           uint32_t* out_lcs, uint64_t* n_pairs)
{
    auto* s = static_cast<SeqSet*>(h);
    std::atomic<int64_t> next((int64_t)row_end - 1);      // big rows first, like CUPGMAQueue
    std::atomic<uint64_t> pairs(0);
    const size_t base = (size_t)row_begin * (row_begin ? row_begin - 1 : 0) / 2;
    // every worker needs its own copy of the row sequence state? No: ComputeBitMasks mutates the
    // row CSequence only, and each row is owned by exactly one worker at a time.
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> workers;
    for (int t = 0; t < n_threads; ++t)
        workers.emplace_back([&] {
            DriverOnly drv(isa_from_int(isa));
            CLCSBP lcsbp(isa_from_int(isa));
            Transform<float, Distance::indel075_div_lcs> tr;
            RawLcs raw;
            std::vector<float> row(row_end ? row_end : 1);
            uint64_t mine = 0;
            for (;;) {
                int64_t i = next.fetch_sub(1);
                if (i < (int64_t)row_begin) break;
                if (out_lcs) {
                    size_t off = (size_t)i * (i ? i - 1 : 0) / 2 - base;
                    drv.calculateDistanceVector<CSequence*, uint32_t, RawLcs>(
                        raw, s->ptrs[i], s->ptrs.data(), (int)i, out_lcs + off, lcsbp);
                } else {
                    drv.calculateDistanceVector<CSequence*, float, decltype(tr)>(
                        tr, s->ptrs[i], s->ptrs.data(), (int)i, row.data(), lcsbp);
                }
                mine += (uint64_t)i;
            }
            pairs += mine;
        });
    for (auto& w : workers) w.join();
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (n_pairs) *n_pairs = pairs.load();
    return sec;
}

// The reference's own Transform functors.  kind: 0 indel075_div_lcs, 1 indel_div_lcs, 2 pairwise_identity.
double ref_transform_f64(int kind, uint32_t lcs, uint32_t len1, uint32_t len2)
{
    if (kind == 0) { Transform<double, Distance::indel075_div_lcs> t; return t(lcs, len1, len2); }
    if (kind == 1) { Transform<double, Distance::indel_div_lcs> t; return t(lcs, len1, len2); }
    Transform<double, Distance::pairwise_identity> t; return t(lcs, len1, len2);
}
float ref_transform_f32(int kind, uint32_t lcs, uint32_t len1, uint32_t len2)
{
    if (kind == 0) { Transform<float, Distance::indel075_div_lcs> t; return t(lcs, len1, len2); }
    if (kind == 1) { Transform<float, Distance::indel_div_lcs> t; return t(lcs, len1, len2); }
    Transform<float, Distance::pairwise_identity> t; return t(lcs, len1, len2);
}

} // extern "C"
