// ref_bridge.cpp — extern "C" doorways into the REFERENCE'S OWN code (compiled unmodified from
// /root/reference by oracle/Makefile into oracle/_ref/libgds_ref.so).  TEST INFRASTRUCTURE.
// Used to pin the oracle: generator streams, coverage helpers, the read_bam pair filter,
// find_pairs, and the reference's CoverageTester driving any solver through qmcp::Solver.
#include <htslib/sam.h>

#include <cstdint>
#include <cstring>
#include <filesystem>
#include <memory>
#include <random>
#include <vector>

#include "bam-api/bam_api.hpp"
#include "bam-api/bam_api_config_builder.hpp"
#include "coverage_tester.hpp"
#include "qmcp-solver/solver.hpp"
#include "reads_gen.hpp"

namespace {
bam_api::AOSPairedReads make_aos(uint64_t n, uint32_t L, const uint32_t* s, const uint32_t* e,
                                 const uint32_t* q, const uint32_t* l) {
    bam_api::AOSPairedReads r;
    r.ref_genome_length = L;
    r.reserve(n);
    for (uint64_t i = 0; i < n; ++i)
        r.push_back(bam_api::Read(i, s[i], e[i], q ? q[i] : 0, l ? l[i] : e[i] - s[i] + 1, i % 2 == 0));
    return r;
}
void dump(const bam_api::AOSPairedReads& r, uint32_t* s, uint32_t* e, uint32_t* q, uint32_t* l) {
    for (size_t i = 0; i < r.reads.size(); ++i) {
        s[i] = (uint32_t)r.reads[i].start_ind;
        e[i] = (uint32_t)r.reads[i].end_ind;
        q[i] = r.reads[i].quality;
        l[i] = r.reads[i].seq_length;
    }
}
}  // namespace

extern "C" {

// libs/reads-gen/src/reads_gen.cpp with the shape lambdas of src/tests/coverage_tester.cpp:157-175
int ref_gen_reads(uint32_t seed, uint64_t pairs, uint32_t L, uint32_t R, int shape,
                  uint32_t* s, uint32_t* e, uint32_t* q, uint32_t* l) {
    std::mt19937 mt(seed);
    bam_api::AOSPairedReads r;
    if (shape == 0) {
        r = reads_gen::rand_reads_uniform(mt, pairs, L, R);
    } else if (shape == 1) {
        r = reads_gen::rand_reads(mt, pairs, L, R, [](double x) { return x - x * x; });
    } else if (shape == 2) {
        r = reads_gen::rand_reads(mt, pairs, L, R, [](double x) {
            if (x > 0.3684 && x < 0.6316) return 1000.0 * (x * x - x + 0.25) * (x * x - x + 0.25) + 0.2;
            return 0.5;
        });
    } else {
        r = reads_gen::rand_reads(mt, pairs, L, R,
                                  [](double x) { return -10.0 * (x - 0.5) * (x - 0.5) + 1.0; });
    }
    dump(r, s, e, q, l);
    return 0;
}

// BamApi::find_input_cover (bam_api.cpp:275-286)
int ref_input_cover(uint64_t n, const uint32_t* s, const uint32_t* e, uint32_t L, uint32_t* cov) {
    bam_api::BamApi api(make_aos(n, L, s, e, nullptr, nullptr));
    auto c = api.find_input_cover();
    memcpy(cov, c.data(), c.size() * sizeof(uint32_t));
    return 0;
}

// BamApi::find_filtered_cover (bam_api.cpp:288-301)
int ref_filtered_cover(uint64_t n, const uint32_t* s, const uint32_t* e, uint32_t L,
                       const uint64_t* ids, uint64_t n_ids, uint32_t* cov) {
    bam_api::BamApi api(make_aos(n, L, s, e, nullptr, nullptr));
    std::vector<bam_api::ReadIndex> v(ids, ids + n_ids);
    auto c = api.find_filtered_cover(v);
    memcpy(cov, c.data(), c.size() * sizeof(uint32_t));
    return 0;
}

// BamApi::find_pairs (bam_api.cpp:239-273)
uint64_t ref_find_pairs(uint64_t n, const uint32_t* s, const uint32_t* e, uint32_t L,
                        const uint64_t* ids, uint64_t n_ids, uint64_t* out) {
    bam_api::BamApi api(make_aos(n, L, s, e, nullptr, nullptr));
    std::vector<bam_api::ReadIndex> v(ids, ids + n_ids);
    auto p = api.find_pairs(v);
    for (size_t i = 0; i < p.size(); ++i) out[i] = p[i];
    return p.size();
}

// The real BamApi::read_bam (bam_api.cpp:359-507) over the fake in-memory BAM: QNAME pair
// matching, should_be_filtered_out (:311-332), set_amplicon_filter from BED/TSV files (:53-187).
// amp_mode: 0 IGNORE (no bed), 1 FILTER.
int64_t ref_read_bam(uint64_t n, const uint32_t* s, const uint32_t* e, const uint32_t* q,
                     const uint32_t* l, uint32_t L, const char* bed_path, const char* tsv_path,
                     uint32_t min_len, uint32_t min_mapq, int amp_mode, uint32_t* os, uint32_t* oe,
                     uint32_t* oq, uint32_t* ol, uint64_t* obam_id, uint64_t* filtered_out,
                     uint64_t* n_filtered_out) {
    gds_fake_bam_set("fake.bam", n, L, s, e, q, l);
    bam_api::BamApiConfigBuilder b;
    b.add_min_mapq(min_mapq);
    b.add_min_seq_length(min_len);
    b.add_hts_thread_count(1);
    if ((amp_mode == 1 || amp_mode == 2) && bed_path && bed_path[0])
        b.add_amplicon_filtering(amp_mode == 1 ? bam_api::AmpliconBehaviour::FILTER
                                               : bam_api::AmpliconBehaviour::GRADE, bed_path,
                                 tsv_path ? std::filesystem::path(tsv_path) : std::filesystem::path());
    bam_api::BamApi api("fake.bam", b.build());
    const bam_api::SOAPairedReads& soa = api.get_paired_reads_soa();
    uint64_t m = soa.get_reads_count();
    for (uint64_t i = 0; i < m; ++i) {
        os[i] = (uint32_t)soa.start_inds[i];
        oe[i] = (uint32_t)soa.end_inds[i];
        oq[i] = soa.qualities[i];
        ol[i] = soa.seq_lengths[i];
        obam_id[i] = soa.ids[i];
    }
    const auto& fo = api.get_filtered_out_reads();
    for (size_t i = 0; i < fo.size(); ++i) filtered_out[i] = fo[i];
    *n_filtered_out = fo.size();
    return (int64_t)m;
}

// Drive ANY solver through the reference's plugin interface and its own test harness.
typedef uint64_t (*gds_solve_cb)(void* user, uint32_t max_coverage, uint64_t n, uint32_t L,
                                 const uint32_t* start, const uint32_t* end, uint64_t* out_ids);

class CallbackSolver : public qmcp::Solver {
   public:
    CallbackSolver(gds_solve_cb cb, void* user) : cb_(cb), user_(user) {}
    std::unique_ptr<qmcp::Solution> solve(uint32_t max_coverage, bam_api::BamApi& api) override {
        const bam_api::SOAPairedReads& soa = api.get_paired_reads_soa();
        uint64_t n = soa.get_reads_count();
        std::vector<uint32_t> s(n), e(n);
        for (uint64_t i = 0; i < n; ++i) {
            s[i] = (uint32_t)soa.start_inds[i];
            e[i] = (uint32_t)soa.end_inds[i];
        }
        std::vector<uint64_t> ids(n);
        uint64_t k = cb_(user_, max_coverage, n, (uint32_t)soa.ref_genome_length, s.data(), e.data(),
                         ids.data());
        auto sol = std::make_unique<qmcp::Solution>(ids.begin(), ids.begin() + k);
        ++calls;
        return sol;
    }
    bool uses_quality_of_reads() override { return false; }
    int calls = 0;

   private:
    gds_solve_cb cb_;
    void* user_;
};

// CoverageTester::test (src/tests/coverage_tester.cpp:28-43): the reference's 5 cases, asserts
// live (built without NDEBUG).  Returns the number of solve() calls (5 on success; a failed
// assert aborts the process, like the reference).
int ref_run_coverage_tests(gds_solve_cb cb, void* user) {
    CallbackSolver solver(cb, user);
    test::CoverageTester tester;
    std::filesystem::path none;
    tester.test(solver, none);
    return solver.calls;
}
}
