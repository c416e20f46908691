// C doorway into the host mirror for Python callers (bench.py, tests): synthetic inputs from
// reads_gen, and the whole plugin path (BamApi + SolverManager + qmcp::Solver) on plain arrays.
// Built into libgds_host.so next to libgds_b200.so.
#include <algorithm>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <filesystem>
#include <random>
#include <stdexcept>
#include <vector>

#include "qmcp-solver/quasi_mcp_b200_max_flow_solver.hpp"
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
// A batch of samples through the plugin: BamApi objects are built from the arrays (what a caller of
// the reference API holds), then QuasiMcpB200MaxFlowSolver::solve_batch — the timed part, reported
// in *solve_seconds — narrows the size_t columns, solves all samples in one device call and expands
// the per-sample index lists.  kept_counts[k] = number of kept reads of sample k; kept_total
// indices are written to kept_out sample after sample (cap permitting).  Returns total kept.
int64_t gdsh_plugin_solve_batch(uint32_t n_samples, const uint64_t* read_off, uint32_t genome_len,
                                const uint32_t* start, const uint32_t* end, uint32_t max_coverage,
                                int repeats, double* solve_seconds, uint64_t* kept_counts,
                                uint64_t* kept_out, uint64_t cap) {
    static qmcp::QuasiMcpB200MaxFlowSolver solver;
    std::vector<std::unique_ptr<bam_api::BamApi>> apis;
    std::vector<bam_api::BamApi*> ptrs;
    for (uint32_t k = 0; k < n_samples; ++k) {
        bam_api::SOAPairedReads soa;
        soa.ref_genome_length = genome_len;
        const uint64_t b = read_off[k], e = read_off[k + 1];
        soa.reserve(e - b);
        for (uint64_t i = b; i < e; ++i)
            soa.push_back(bam_api::Read(i - b, start[i], end[i], 0, end[i] - start[i] + 1, (i - b) % 2 == 0));
        apis.push_back(std::make_unique<bam_api::BamApi>(soa));
        ptrs.push_back(apis.back().get());
    }
    std::vector<std::unique_ptr<qmcp::Solution>> sols;
    double best = 1e30;
    for (int r = 0; r < (repeats > 0 ? repeats : 1); ++r) {
        auto t0 = std::chrono::steady_clock::now();
        sols = solver.solve_batch(max_coverage, ptrs);
        best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    if (solve_seconds) *solve_seconds = best;
    uint64_t total = 0;
    for (uint32_t k = 0; k < n_samples; ++k) {
        if (kept_counts) kept_counts[k] = sols[k]->size();
        for (uint64_t v : *sols[k]) {
            if (kept_out && total < cap) kept_out[total] = v;
            ++total;
        }
    }
    return static_cast<int64_t>(total);
}

}  // extern "C"

// The narrowing a caller with 32-bit columns has to do to use the compact transport
// (include/gds.h gds_reads.start16 / end == NULL): 16-bit starts, on `threads` host threads.
// end != NULL: the exact read-length range is computed as well (what the C++ adapter's narrowing loop
// does with the reference's size_t columns).  end == NULL: the caller KNOWS the one read length (the
// generator's read_length parameter, a run's metadata) and passes it as the hint; only the start
// column is read.  Returns 1 when every start fits 16 bits.  AVX2 when the CPU has it: the loop is
// memory-bound either way, but 8 threads of scalar code do not reach the memory system's rate.
namespace {
struct EncPart {
    uint32_t lo = 0xffffffffu, hi = 0, big = 0;
};

#if defined(__x86_64__)
__attribute__((target("avx2"))) void enc_range_avx2(const uint32_t* start, const uint32_t* end,
                                                    uint16_t* out, uint64_t b, uint64_t e, EncPart& p) {
    uint64_t i = b;
    __m256i vbig = _mm256_setzero_si256();
    __m256i vlo = _mm256_set1_epi32(-1), vhi = _mm256_setzero_si256();
    // the 16-bit column is read next by the copy engine, not by this core: once the output is
    // 32-byte aligned it is written with streaming stores (no read-for-ownership of 2 bytes per
    // read, nothing evicted from the caches)
    for (; i < e && (reinterpret_cast<uintptr_t>(out + i) & 31u); ++i) {
        const uint32_t s = start[i];
        out[i] = static_cast<uint16_t>(s);
        p.big |= s;
        if (end) {
            const uint32_t len = end[i] - s + 1;
            p.lo = std::min(p.lo, len);
            p.hi = std::max(p.hi, len);
        }
    }
    for (; i + 16 <= e; i += 16) {
        const __m256i s0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(start + i));
        const __m256i s1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(start + i + 8));
        vbig = _mm256_or_si256(vbig, _mm256_or_si256(s0, s1));
        // packus saturates, which is fine: a start beyond 16 bits fails the call anyway
        const __m256i pk = _mm256_permute4x64_epi64(_mm256_packus_epi32(s0, s1), 0xD8);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(out + i), pk);
        if (end) {
            const __m256i one = _mm256_set1_epi32(1);
            const __m256i l0 = _mm256_add_epi32(
                _mm256_sub_epi32(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(end + i)), s0), one);
            const __m256i l1 = _mm256_add_epi32(
                _mm256_sub_epi32(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(end + i + 8)), s1), one);
            vlo = _mm256_min_epu32(vlo, _mm256_min_epu32(l0, l1));
            vhi = _mm256_max_epu32(vhi, _mm256_max_epu32(l0, l1));
        }
    }
    _mm_sfence();
    alignas(32) uint32_t t[8];
    _mm256_store_si256(reinterpret_cast<__m256i*>(t), vbig);
    for (uint32_t x : t) p.big |= x;
    if (end) {
        _mm256_store_si256(reinterpret_cast<__m256i*>(t), vlo);
        for (uint32_t x : t) p.lo = std::min(p.lo, x);
        _mm256_store_si256(reinterpret_cast<__m256i*>(t), vhi);
        for (uint32_t x : t) p.hi = std::max(p.hi, x);
    }
    for (; i < e; ++i) {
        const uint32_t s = start[i];
        out[i] = static_cast<uint16_t>(s);
        p.big |= s;
        if (end) {
            const uint32_t len = end[i] - s + 1;
            p.lo = std::min(p.lo, len);
            p.hi = std::max(p.hi, len);
        }
    }
}
#endif

void enc_range(const uint32_t* start, const uint32_t* end, uint16_t* out, uint64_t b, uint64_t e,
               EncPart& p) {
#if defined(__x86_64__)
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) {
        enc_range_avx2(start, end, out, b, e, p);
        return;
    }
#endif
    for (uint64_t i = b; i < e; ++i) {
        const uint32_t s = start[i];
        out[i] = static_cast<uint16_t>(s);
        p.big |= s;
        if (end) {
            const uint32_t len = end[i] - s + 1;
            p.lo = std::min(p.lo, len);
            p.hi = std::max(p.hi, len);
        }
    }
}
}  // namespace

extern "C" int gdsh_encode_compact(const uint32_t* start, const uint32_t* end, uint64_t n, uint16_t* start16,
                                   uint32_t* len_min, uint32_t* len_max, uint32_t threads) {
    const unsigned nt = n < (1u << 18) ? 1u : std::max(1u, std::min(threads, 64u));
    std::vector<EncPart> parts(nt);
    const uint64_t per = ((n + nt - 1) / nt + 63) & ~uint64_t{63};
    auto work = [&](unsigned t) {
        const uint64_t b = std::min<uint64_t>(n, (uint64_t)t * per), e = std::min<uint64_t>(n, b + per);
        enc_range(start, end, start16, b, e, parts[t]);
    };
    if (nt == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    int ok = 1;
    *len_min = 0xffffffffu;
    *len_max = 0;
    for (const EncPart& p : parts) {
        *len_min = std::min(*len_min, p.lo);
        *len_max = std::max(*len_max, p.hi);
        ok &= p.big <= 0xffffu;
    }
    return ok;
}

extern "C" {
// ---- BAM files: the file-backed BamApi behind a handle ----

int64_t gdsh_write_synthetic_bam(const char* path, uint64_t n, uint32_t genome_len,
                                 const uint32_t* start, const uint32_t* end, const uint8_t* mapq,
                                 const uint32_t* seq_len, int coordinate_sorted, uint32_t threads,
                                 uint32_t seed) {
    try {
        return static_cast<int64_t>(reads_gen::write_synthetic_bam(path, n, genome_len, start, end, mapq,
                                                                   seq_len, coordinate_sorted != 0, threads, seed));
    } catch (const std::exception&) {
        return -1;
    }
}

// BamApi(input_filepath, config) as App builds it (src/app.cpp:113-128); bed/tsv may be empty.
// amplicon_behaviour: 0 IGNORE, 1 FILTER, 2 GRADE
void* gdsh_bam_open(const char* path, const char* bed, const char* tsv, uint32_t min_len,
                    uint32_t min_mapq, int amplicon_behaviour, uint32_t threads) {
    bam_api::BamApiConfig cfg;
    cfg.bed_filepath = bed ? bed : "";
    cfg.tsv_filepath = tsv ? tsv : "";
    cfg.min_seq_length = min_len;
    cfg.min_mapq = min_mapq;
    cfg.hts_thread_count = threads;
    cfg.amplicon_behaviour = static_cast<bam_api::AmpliconBehaviour>(amplicon_behaviour);
    return new bam_api::BamApi(std::filesystem::path(path), cfg);
}

void gdsh_bam_close(void* h) { delete static_cast<bam_api::BamApi*>(h); }

static uint64_t copy_soa(const bam_api::SOAPairedReads& soa, uint64_t cap, uint64_t* ids,
                         uint64_t* start, uint64_t* end, uint32_t* quality, uint32_t* seq_len,
                         uint8_t* is_first) {
    uint64_t n = soa.get_reads_count(), k = n < cap ? n : cap;
    for (uint64_t i = 0; i < k; ++i) {
        if (ids) ids[i] = soa.ids[i];
        if (start) start[i] = soa.start_inds[i];
        if (end) end[i] = soa.end_inds[i];
        if (quality) quality[i] = soa.qualities[i];
        if (seq_len) seq_len[i] = soa.seq_lengths[i];
        if (is_first) is_first[i] = soa.is_first_reads[i];
    }
    return n;
}

// the pair-ordered reads BEFORE the filter (reads the file on first use); returns their count
uint64_t gdsh_bam_unfiltered(void* h, uint64_t cap, uint64_t* ids, uint64_t* start, uint64_t* end,
                             uint32_t* quality, uint32_t* seq_len, uint8_t* is_first) {
    auto* api = static_cast<bam_api::BamApi*>(h);
    if (!api->has_pending_filter()) return 0;
    return copy_soa(api->unfiltered_soa(), cap, ids, start, end, quality, seq_len, is_first);
}

// get_paired_reads_soa(): the state read_bam leaves (host filter unless a device solve ran first)
uint64_t gdsh_bam_reads(void* h, uint64_t cap, uint64_t* ids, uint64_t* start, uint64_t* end,
                        uint32_t* quality, uint32_t* seq_len, uint8_t* is_first) {
    auto* api = static_cast<bam_api::BamApi*>(h);
    return copy_soa(api->get_paired_reads_soa(), cap, ids, start, end, quality, seq_len, is_first);
}

uint64_t gdsh_bam_filtered_out(void* h, uint64_t cap, uint64_t* out) {
    const auto& v = static_cast<bam_api::BamApi*>(h)->get_filtered_out_reads();
    std::memcpy(out, v.data(), (v.size() < cap ? v.size() : cap) * sizeof(uint64_t));
    return v.size();
}

uint64_t gdsh_bam_ref_length(void* h) {
    auto* api = static_cast<bam_api::BamApi*>(h);
    api->has_pending_filter();
    return api->get_paired_reads().ref_genome_length;
}

uint64_t gdsh_bam_record_count(void* h) {
    auto* api = static_cast<bam_api::BamApi*>(h);
    api->has_pending_filter();
    return api->bam_record_count();
}

double gdsh_bam_read_seconds(void* h) { return static_cast<bam_api::BamApi*>(h)->read_bam_seconds(); }

// solve through the plugin interface on the file-backed BamApi (src/app.cpp:130-135)
int64_t gdsh_bam_solve(void* h, const char* algorithm, uint32_t max_coverage, uint64_t* kept_out,
                       uint64_t cap) {
    static SolverManager manager;
    if (!manager.contains(algorithm)) return -1;
    auto sol = manager.get(algorithm).solve(max_coverage, *static_cast<bam_api::BamApi*>(h));
    uint64_t k = sol->size() < cap ? sol->size() : cap;
    std::memcpy(kept_out, sol->data(), k * sizeof(uint64_t));
    return static_cast<int64_t>(sol->size());
}

// find_pairs + write_paired_reads as App::execute chains them (src/app.cpp:141-147)
int64_t gdsh_bam_write_solution(void* h, const char* out_path, const uint64_t* kept, uint64_t n,
                                int with_pairs) {
    auto* api = static_cast<bam_api::BamApi*>(h);
    std::vector<bam_api::ReadIndex> ids(kept, kept + n);
    if (with_pairs) ids = api->find_pairs(ids);
    return api->write_paired_reads(out_path, ids);
}

int64_t gdsh_bam_write_filtered_out(void* h, const char* out_path) {
    return static_cast<bam_api::BamApi*>(h)->write_bam_api_filtered_out_reads(out_path);
}

int64_t gdsh_write_bam(const char* in_path, const char* out_path, const uint64_t* bam_ids, uint64_t n,
                       uint32_t threads) {
    std::vector<bam_api::BAMReadId> ids(bam_ids, bam_ids + n);
    return bam_api::BamApi::write_bam(in_path, out_path, ids, threads);
}
}
