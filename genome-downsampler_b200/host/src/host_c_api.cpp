// C doorway into the host mirror for Python callers (bench.py, tests): synthetic inputs from
// reads_gen, and the whole plugin path (BamApi + SolverManager + qmcp::Solver) on plain arrays.
// Built into libgds_host.so next to libgds_b200.so.
#include <cstdint>
#include <cstring>
#include <random>

#include "reads_gen.hpp"
#include "solver_manager.hpp"

extern "C" {

// shape: 0 = rand_reads_uniform; 1..3 = the three histogram shapes the reference tests with
// (src/tests/coverage_tester.cpp:157-175)
int gdsh_gen_reads(uint32_t seed, uint64_t pairs, uint32_t genome_len, uint32_t read_len, int shape,
                   uint32_t* start, uint32_t* end, uint8_t* mapq, uint32_t* seq_len) {
    if (genome_len < 2 * read_len || read_len == 0) return 1;
    std::mt19937 mt(seed);
    reads_gen::SoaOut out{start, end, mapq, seq_len};
    switch (shape) {
        case 0:
            reads_gen::rand_reads_uniform_soa(mt, pairs, genome_len, read_len, out);
            break;
        case 1:
            reads_gen::rand_reads_soa(mt, pairs, genome_len, read_len,
                                      [](double x) { return x - x * x; }, out);
            break;
        case 2:
            reads_gen::rand_reads_soa(mt, pairs, genome_len, read_len, [](double x) {
                double c = x * x - x + 0.25;
                return (x > 0.3684 && x < 0.6316) ? 1000.0 * c * c + 0.2 : 0.5;
            }, out);
            break;
        case 3:
            reads_gen::rand_reads_soa(mt, pairs, genome_len, read_len,
                                      [](double x) { return 1.0 - 10.0 * (x - 0.5) * (x - 0.5); }, out);
            break;
        default:
            return 2;
    }
    return 0;
}

// config-2 style reads: mates inside a random amplicon with probability p_inside
int gdsh_gen_reads_amplicon(uint32_t seed, uint64_t pairs, uint32_t genome_len, uint32_t n_amp,
                            const uint32_t* amp_start, const uint32_t* amp_end, double p_inside,
                            uint32_t min_len, uint32_t max_len, uint32_t* start, uint32_t* end,
                            uint8_t* mapq, uint32_t* seq_len) {
    if (n_amp == 0 || max_len < min_len || genome_len < 2 * max_len) return 1;
    std::mt19937 mt(seed);
    std::vector<uint32_t> a0(amp_start, amp_start + n_amp), a1(amp_end, amp_end + n_amp);
    reads_gen::rand_reads_amplicon_soa(mt, pairs, genome_len, a0, a1, p_inside, min_len, max_len,
                                       reads_gen::SoaOut{start, end, mapq, seq_len});
    return 0;
}

// One downsample through the plugin interface exactly as App::execute drives it
// (src/app.cpp:130-135): SolverManager.get(name).solve(max_coverage, bam_api).
// Returns the number of kept indices written (ascending), or -1 for an unknown algorithm.
int64_t gdsh_plugin_solve(const char* algorithm, uint64_t n, uint32_t genome_len,
                          const uint32_t* start, const uint32_t* end, uint32_t max_coverage,
                          uint64_t* kept_out, uint64_t cap) {
    static SolverManager manager;  // solvers live for the process, like the reference's registry
    if (!manager.contains(algorithm)) return -1;
    bam_api::SOAPairedReads soa;
    soa.ref_genome_length = genome_len;
    soa.reserve(n);
    for (uint64_t i = 0; i < n; ++i)
        soa.push_back(bam_api::Read(i, start[i], end[i], 0, end[i] - start[i] + 1, i % 2 == 0));
    bam_api::BamApi api(soa);
    auto sol = manager.get(algorithm).solve(max_coverage, api);
    uint64_t k = sol->size() < cap ? sol->size() : cap;
    std::memcpy(kept_out, sol->data(), k * sizeof(uint64_t));
    return static_cast<int64_t>(sol->size());
}
}
