// gds_host_test — the reference's `test` subcommand for this path (src/test_command.cpp:31-70,
// src/tests/coverage_tester.cpp): runs the "coverage" tester's five cases against the solvers
// selected with -a through the qmcp::Solver interface, with LIVE checks (the reference's asserts
// vanish under NDEBUG), plus a device-filter case (config 2 shape).
//   gds_host_test [-a quasi-mcp-b200] [-o DIR] [-b N_SAMPLES]   (-b: solve_batch throughput)
#include <chrono>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <random>
#include <string>

#include <memory>

#include "logging/log.hpp"
#include "qmcp-solver/quasi_mcp_b200_max_flow_solver.hpp"
#include "reads_gen.hpp"
#include "solver_manager.hpp"

namespace fs = std::filesystem;
using Cover = std::vector<uint32_t>;

static bool is_out_cover_valid(const Cover& in, const Cover& out, uint32_t m) {
    // coverage_tester.cpp:101-107: min(in, m) <= out elementwise
    for (size_t i = 0; i < in.size(); ++i)
        if (std::min(in[i], m) > out[i]) return false;
    return true;
}

static bam_api::AOSPairedReads small_example() {
    // the 16 reads of coverage_tester.cpp:72-93 as (start, end)
    static const uint32_t se[16][2] = {{0, 2}, {6, 9}, {2, 4}, {6, 8}, {1, 3}, {7, 10}, {3, 6}, {9, 10},
                                       {0, 4}, {7, 9}, {4, 6}, {9, 10}, {1, 4}, {6, 8}, {0, 2}, {4, 6}};
    bam_api::AOSPairedReads r;
    r.ref_genome_length = 11;
    for (uint32_t i = 0; i < 16; ++i)
        r.push_back(bam_api::Read(i, se[i][0], se[i][1], 0, se[i][1] - se[i][0] + 1, i % 2 == 0));
    return r;
}

struct Case {
    std::string name;
    uint32_t m;
    std::function<bam_api::AOSPairedReads()> make;
};

static int run_case(qmcp::Solver& solver, const Case& c, const fs::path& outdir) {
    auto input = c.make();
    bam_api::BamApi api(input);
    Cover in_cover = api.find_input_cover();
    // wall time of solve() as the reference logs it (src/tests/scoped_timer.hpp:9-13, app.cpp:132-139)
    auto t0 = std::chrono::steady_clock::now();
    auto ids = solver.solve(c.m, api);
    double wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    double dev_ms = -1;
    if (auto* b200 = dynamic_cast<qmcp::QuasiMcpB200MaxFlowSolver*>(&solver))
        dev_ms = b200->last_result().ms_total;
    Cover out_cover = api.find_filtered_cover(*ids);
    bool ok = is_out_cover_valid(in_cover, out_cover, c.m);
    bool sorted = std::is_sorted(ids->begin(), ids->end()) &&
                  std::adjacent_find(ids->begin(), ids->end()) == ids->end();
    LOG_WITH_LEVEL(logging::INFO) << "  " << c.name << ": reads=" << input.reads.size()
                                  << " kept=" << ids->size() << " solve() took " << wall_ms
                                  << " ms (device " << dev_ms << " ms)"
                                  << (ok && sorted ? " PASSED" : " FAILED");
    if (!outdir.empty()) {
        std::ofstream f(outdir / (c.name + ".cov"));
        for (size_t i = 0; i < in_cover.size(); ++i) f << i << "\t" << in_cover[i] << "\t" << out_cover[i] << "\n";
    }
    return ok && sorted ? 0 : 1;
}

static int run_filter_case(qmcp::Solver& solver) {
    // unfiltered reads + min length / MAPQ: device filter result must equal the host predicate
    std::mt19937 mt(2024);
    auto aos = reads_gen::rand_reads_uniform(mt, 20000, 30000, 150);
    std::uniform_int_distribution<> len(60, 150);
    for (auto& r : aos.reads) r.seq_length = len(mt);
    bam_api::SOAPairedReads soa;
    soa.from(aos);
    bam_api::BamApiConfig cfg;
    cfg.min_seq_length = 90;
    cfg.min_mapq = 30;
    bam_api::BamApi dev_api(soa, cfg), host_api(soa, cfg);
    auto ids = solver.solve(50, dev_api);
    const auto& want = host_api.get_paired_reads_soa();  // host filter path
    const auto& got = dev_api.get_paired_reads_soa();
    bool same = want.ids == got.ids && dev_api.get_filtered_out_reads() == host_api.get_filtered_out_reads();
    Cover in_cover = dev_api.find_input_cover();
    Cover out_cover = dev_api.find_filtered_cover(*ids);
    bool ok = same && is_out_cover_valid(in_cover, out_cover, 50);
    LOG_WITH_LEVEL(logging::INFO) << "  device_filter_l90_q30: survivors=" << got.get_reads_count()
                                  << " kept=" << ids->size() << (ok ? " PASSED" : " FAILED");
    return ok ? 0 : 1;
}

// A batch of independent samples through QuasiMcpB200MaxFlowSolver::solve_batch (BASELINE config 5
// shape: 2 M reads over 30 kb each, M = 100): wall time of the call — narrowing of the reference's
// size_t columns, one device call, per-sample index lists — as reads/s, and each sample's result
// against solve() on the same BamApi.
static int run_batch(uint32_t n_samples) {
    qmcp::QuasiMcpB200MaxFlowSolver solver;
    std::vector<std::unique_ptr<bam_api::BamApi>> apis;
    std::vector<bam_api::BamApi*> ptrs;
    uint64_t reads = 0;
    for (uint32_t k = 0; k < n_samples; ++k) {
        std::mt19937 mt(12345 + k);
        auto aos = reads_gen::rand_reads_uniform(mt, 1'000'000, 30'000, 150);
        reads += aos.reads.size();
        apis.push_back(std::make_unique<bam_api::BamApi>(aos));
        apis.back()->get_paired_reads_soa();  // the AoS -> SoA conversion is the caller's, not the solver's
        ptrs.push_back(apis.back().get());
    }
    double best = 1e30;
    std::vector<std::unique_ptr<qmcp::Solution>> sols;
    for (int rep = 0; rep < 3; ++rep) {
        auto t0 = std::chrono::steady_clock::now();
        sols = solver.solve_batch(100, ptrs);
        best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    bool ok = sols.size() == n_samples;
    for (uint32_t k : {0u, n_samples / 2, n_samples - 1}) {
        auto one = solver.solve(100, *ptrs[k]);
        ok = ok && *one == *sols[k];
    }
    LOG_WITH_LEVEL(logging::INFO) << "  solve_batch: " << n_samples << " samples, " << reads << " reads in "
                                  << best * 1e3 << " ms = " << reads / best / 1e9 << " G reads/s (device "
                                  << solver.last_result().ms_total << " ms)" << (ok ? " PASSED" : " FAILED");
    std::printf("solve_batch %u samples %.3f ms %.3f Greads/s\n", n_samples, best * 1e3, reads / best / 1e9);
    return ok ? 0 : 1;
}

int main(int argc, char** argv) {
    SolverManager manager;
    std::vector<std::string> algs;
    fs::path outdir;
    uint32_t batch = 0;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "-b") && i + 1 < argc) {
            batch = static_cast<uint32_t>(std::atoi(argv[++i]));
            continue;
        }
        if (!strcmp(argv[i], "-a") && i + 1 < argc) algs.push_back(argv[++i]);
        else if (!strcmp(argv[i], "-o") && i + 1 < argc) outdir = argv[++i];
        else if (!strcmp(argv[i], "-v")) SET_LOG_LEVEL(logging::DEBUG);
    }
    if (algs.empty()) algs = manager.get_names();
    auto shaped = [](std::function<double(double)> f) {
        return [f]() {
            std::mt19937 mt(12345);
            return reads_gen::rand_reads(mt, 1'000'000, 30'000, 150, f);
        };
    };
    std::vector<Case> cases = {
        {"small_example_test", 4, small_example},
        {"random_uniform_dist_test", 1000,
         []() {
             std::mt19937 mt(12345);
             return reads_gen::rand_reads_uniform(mt, 1'000'000, 30'000, 150);
         }},
        {"random_low_coverage_on_both_sides_test", 8000, shaped([](double x) { return x - x * x; })},
        {"random_with_hole_test", 8000, shaped([](double x) {
             if (x > 0.3684 && x < 0.6316) return 1000.0 * (x * x - x + 0.25) * (x * x - x + 0.25) + 0.2;
             return 0.5;
         })},
        {"random_zero_coverage_on_both_sides_test", 8000,
         shaped([](double x) { return -10.0 * (x - 0.5) * (x - 0.5) + 1.0; })},
    };
    int failures = 0;
    for (const auto& a : algs) {
        if (!manager.contains(a)) {
            LOG_WITH_LEVEL(logging::ERROR) << "unknown algorithm " << a;
            return 2;
        }
        LOG_WITH_LEVEL(logging::INFO) << "Running test coverage on algorithm " << a;
        for (const auto& c : cases) failures += run_case(manager.get(a), c, outdir);
        failures += run_filter_case(manager.get(a));
    }
    if (batch) failures += run_batch(batch);
    std::printf("%s\n", failures ? "FAILED" : "ALL PASSED");
    return failures ? 1 : 0;
}
