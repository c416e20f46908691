/*
 * gds_oracle.h — CPU ORACLE for the quasi-MCP hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product path (genome-downsampler_b200/, include/gds.h) may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, as the checker / the timed CPU baseline.
 *
 * PARITY PIN STATUS: the reference ships no golden vectors for this path and its max-flow
 * arithmetic lives in Google OR-Tools v9.9.3963 (scripts/install_libs.sh:109-111), which is absent
 * here.  What pins this oracle:
 *   - the reference's own reads_gen / container / filter sources compiled unmodified into
 *     oracle/_ref (see oracle/Makefile) — generator streams, find_input_cover and the pair filter
 *     predicate are compared bit-for-bit in tests/test_oracle_vs_ref.py (run here, fixtures
 *     committed under tests/golden/);
 *   - the reference's only result assertion, CoverageTester::is_out_cover_valid
 *     (src/tests/coverage_tester.cpp:101-107), on its 5 cases;
 *   - known answers for the 16-read example and scipy's Dinic for F*.
 * The kept-read SET of OR-Tools is not reproducible (any max flow is accepted by the reference),
 * so for the kept set the status is "parity unpinned": the deterministic schedule below is
 * self-defined and the GPU must match THIS restatement bit-exactly.
 */
#ifndef GDS_ORACLE_H
#define GDS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- generators: libs/reads-gen/src/reads_gen.cpp:5-86 ---- */
/* shape: 0 uniform (rand_reads_uniform), 1 x-x^2, 2 hole, 3 zero flanks (coverage_tester.cpp:157-175) */
int orc_gen_reads(uint32_t seed, uint64_t pairs, uint32_t genome_len, uint32_t read_len, int shape,
                  int32_t max_quality, uint32_t* start, uint32_t* end, uint32_t* quality,
                  uint32_t* seq_len);
/* amplicon-aware extension for config 2; writes BED/TSV text into caller buffers */
int orc_gen_artic_scheme(uint32_t genome_len, uint32_t n_amplicons, uint32_t amp_len,
                         uint32_t overlap, uint32_t primer_len, char* bed, uint64_t bed_cap,
                         char* tsv, uint64_t tsv_cap);
int orc_gen_reads_amplicon(uint32_t seed, uint64_t pairs, uint32_t genome_len, uint32_t n_amp,
                           const uint32_t* amp_start, const uint32_t* amp_end, double p_inside,
                           uint32_t min_len, uint32_t max_len, int32_t max_quality, uint32_t* start,
                           uint32_t* end, uint32_t* quality, uint32_t* seq_len);

/* ---- amplicon table: libs/bam-api/src/bam_api.cpp:53-95,101-187 ---- */
int64_t orc_parse_amplicons(const char* bed_text, const char* tsv_text /* may be NULL */,
                            uint32_t* amp_start, uint32_t* amp_end, uint64_t cap);

/* ---- filter: bam_api.cpp:311-332, amplicon.cpp:5-7, amplicon_set.cpp:5-9 ---- */
/* pair_pass[p] = 1 iff pair (2p, 2p+1) is kept.  returns number of reads kept. */
uint64_t orc_filter_pairs(uint64_t n_reads, const uint32_t* start, const uint32_t* end,
                          const uint32_t* quality, const uint32_t* seq_len, uint32_t min_len,
                          uint32_t min_mapq, int use_amplicons, uint32_t n_amp,
                          const uint32_t* amp_start, const uint32_t* amp_end, uint8_t* pair_pass);

/* ---- coverage / demand: quasi_mcp_cpu_max_flow_solver.cpp:58-87, bam_api.cpp:275-301 ---- */
/* per-base increments exactly as the reference does it (O(sum of read lengths)) */
int orc_coverage_ref(uint64_t n, const uint32_t* start, const uint32_t* end, uint32_t L,
                     uint32_t* cov /* L */);
/* subset variant (find_filtered_cover): kept is a byte mask */
int orc_coverage_subset(uint64_t n, const uint32_t* start, const uint32_t* end, const uint8_t* kept,
                        uint32_t L, uint32_t* cov);
/* demand[0..L] from cov[0..L-1]; returns F* = sum of source capacities */
int64_t orc_demand(const uint32_t* cov, uint32_t L, uint32_t M, int32_t* demand /* L+1 */);

/* ---- quasi-mcp-cpu restatement: quasi_mcp_cpu_max_flow_solver.cpp:11-100 ---- */
typedef struct {
    int64_t flow_value;
    uint64_t n_kept;
    uint64_t n_arcs;
    uint64_t pushes, relabels, global_relabels;
    double t_coverage_s, t_graph_s, t_maxflow_s, t_select_s, t_total_s;
} orc_ref_stats;
/* kept[i] = 1 iff Flow(arc i) > 0.  Sequential push-relabel stands in for
 * operations_research::SimpleMaxFlow::Solve (OR-Tools absent). */
int orc_ref_solve(uint64_t n, const uint32_t* start, const uint32_t* end, uint32_t L, uint32_t M,
                  uint8_t* kept, orc_ref_stats* st);

/* ---- deterministic bulk-synchronous schedule on the bundled graph ---- */
typedef struct {
    uint32_t gr_interval_min; /* minimum rounds between global relabels */
    uint32_t gr_levels_pct;   /* ... and at least this % of the last BFS depth */
    uint32_t gr_relabel_pct;  /* trigger when relabels since last GR >= pct% of component nodes */
    uint32_t max_rounds;      /* safety stop (0 = unlimited) */
    uint32_t seg_len;         /* segment length: references longer than 2*seg_len are cut (0 = default rule) */
    uint32_t schedule;        /* 0 = express schedule where eligible (gds_params.schedule), 1 = classic only,
                                 2 = express for every component that is structurally eligible (lab),
                                 3 = the graph reduction at every M, express by the usual rule */
} orc_sync_params;
typedef struct {
    int64_t flow_value; /* total sink inflow over all components */
    int64_t fstar;      /* closed form */
    uint64_t n_kept;
    uint64_t n_bundles;
    uint32_t n_components;
    uint64_t rounds_total, rounds_max; /* summed / max over components */
    uint64_t pushes, relabels, global_relabels, bfs_levels;
    uint64_t max_frontier;
    uint32_t n_express; /* components that ran the express schedule */
    double t_build_s, t_solve_s, t_select_s;
} orc_sync_stats;
/* Batch-aware: sample k owns reads [read_off[k], read_off[k+1]) and positions 0..ref_len[k]-1.
 * kept_bitmap has ceil(n_total/32) words, bit i = read i of the concatenated arrays.
 * demand_out (optional) has sum(ref_len[k]+1) entries, cov_out likewise (covR per node). */
int orc_sync_solve(uint32_t n_samples, const uint64_t* read_off, const uint32_t* ref_len,
                   const uint32_t* start, const uint32_t* end, uint32_t M,
                   const orc_sync_params* prm, uint32_t* kept_bitmap, int32_t* demand_out,
                   uint32_t* cov_out, orc_sync_stats* st);

/* The same inputs and outputs for the minimum-cardinality solve (mcp-cpu's objective,
 * mcp_cpu_cost_scaling_solver.cpp:33-67) as the deterministic sweep of csrc/sweep.cuh. */
int orc_sweep_solve(uint32_t n_samples, const uint64_t* read_off, const uint32_t* ref_len,
                    const uint32_t* start, const uint32_t* end, uint32_t M,
                    const orc_sync_params* prm, uint32_t* kept_bitmap, int32_t* demand_out,
                    uint32_t* cov_out, orc_sync_stats* st);

/* ---- min-cardinality optimum (objective of mcp-cpu, mcp_cpu_cost_scaling_solver.cpp:33-67) ---- */
/* greedy interval multicover; returns number of reads kept, kept[] byte mask */
uint64_t orc_greedy_multicover(uint64_t n, const uint32_t* start, const uint32_t* end, uint32_t L,
                               uint32_t M, uint8_t* kept);

/* ---- find_pairs: bam_api.cpp:239-273 (bitmap form: kept |= mate kept) ---- */
void orc_find_pairs_bitmap(uint64_t n, uint32_t* bitmap);

#ifdef __cplusplus
}
#endif
#endif
