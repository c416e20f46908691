// maxflow.cuh — K3: bulk-synchronous push-relabel with on-device global relabelling.
//
// Replaces push_relabel_kernel + host global_relabel + driver loop of the reference
// (quasi_mcp_cuda_max_flow_solver.cu:12-79, :101-155, :366-405) with a different design:
//   * one CTA per connected component (K4), persistent over a work counter; all state stays in
//     HBM/L2 for the whole solve, nothing goes back to the host between rounds;
//   * frontier queues staged in shared memory (first kQCap entries) with a global spill slice;
//     appends are warp-aggregated (one shared-memory atomic per coalesced group);
//   * every round is a pure function of the previous state: phase A pushes along
//     admissible arcs using the labels of the round start, received excess is accumulated with
//     commutative atomics in a side array; phase B merges it, relabels from a label SNAPSHOT and
//     builds the next frontier.  The schedule is therefore independent of thread timing and the
//     CPU oracle (oracle/gds_oracle.cpp: sync_solve_component) replays it bit-exactly;
//   * global relabel = level-synchronous reverse BFS from the sink inside the same CTA, triggered
//     by a deterministic state-only rule;
//   * source arcs are consumed by the preflow (initial excess), the sink is implicit (snk[v]).
//     On this network every active node always has a residual path to the sink,
//     so no excess ever has to return to the source.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "graph.cuh"

namespace gds {
namespace cg = cooperative_groups;

// Launch shapes of the same kernel (the schedule does not depend on the thread count).  Every
// round is latency-bound with a small frontier, so throughput comes from running ALL components
// concurrently: the host picks the widest CTA whose resident count (CTAs/SM x 148) still covers
// the number of components; beyond 8 CTAs/SM the work counter hands out the rest in waves.
//   1024 threads x 1/SM (4096-entry staged queues) ... 128 threads x 8/SM (1024-entry queues)
struct MfShape {
    int threads, ctas_per_sm;
    uint32_t qcap;
};
constexpr MfShape kMfShapes[4] = {{1024, 1, 4096}, {512, 2, 2048}, {256, 4, 1024}, {128, 8, 1024}};

struct BundleGraph {
    const uint32_t* b_s;
    const uint32_t* b_t;
    const uint32_t* b_mult;
    uint32_t* f;  // bundle flows
    const uint32_t* out_ptr;
    const uint32_t* in_ptr;
    const uint32_t* in_bid;
};

struct SolveParams {
    uint32_t gr_interval_min, gr_levels_pct, gr_relabel_pct, max_rounds;
};

struct CompStats {
    unsigned long long rounds, pushes, relabels, grs, bfs_levels, max_frontier;
    long long sink_flow, stuck;
    unsigned long long cycles, frontier_sum;  // diagnostics: SM clocks spent, sum of frontier sizes
};

template <uint32_t QCAP>
struct Queue {
    uint32_t* sm;  // QCAP entries in shared memory
    uint32_t* gl;  // global spill (indexed by the same position)
    __device__ __forceinline__ uint32_t get(uint32_t i) const { return i < QCAP ? sm[i] : gl[i]; }
    __device__ __forceinline__ void put(uint32_t i, uint32_t v) const {
        if (i < QCAP) sm[i] = v;
        else gl[i] = v;
    }
};

// warp-aggregated append from divergent code
template <uint32_t QCAP>
__device__ __forceinline__ void q_append(const Queue<QCAP>& q, uint32_t* count, uint32_t v) {
    auto g = cg::coalesced_threads();
    uint32_t base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(count, g.size());
    base = g.shfl(base, 0);
    q.put(base + g.thread_rank(), v);
}

template <uint32_t QCAP>
struct MfShared {
    uint32_t qa[QCAP], qb[QCAP], qc[QCAP];
    uint32_t nF, nT, nN;
    uint32_t relabels_since;
    uint32_t comp;
    unsigned long long pushes, relabels;
    long long sink_flow, stuck;
};

// reverse BFS from the sink; T/N are used as the level queues.  returns the level counter
template <int THREADS, uint32_t QCAP>
__device__ uint32_t mf_global_relabel(const NodeArrays& na, const BundleGraph& bg, uint32_t lo,
                                      uint32_t hi, Queue<QCAP> T, Queue<QCAP> N,
                                      MfShared<QCAP>& sh, unsigned long long& bfs_levels) {
    constexpr uint32_t kMfThreads = THREADS;
    const uint32_t tid = threadIdx.x;
    for (uint32_t v = lo + tid; v <= hi; v += kMfThreads) na.d_cur[v] = kLabelInf;
    if (tid == 0) {
        sh.nT = 0;
        sh.nN = 0;
    }
    __syncthreads();
    for (uint32_t v = lo + tid; v <= hi; v += kMfThreads) {
        if (na.snk[v] > 0) {
            na.d_cur[v] = 1;
            q_append(T, &sh.nT, v);
        }
    }
    __syncthreads();
    uint32_t level = 1;
    uint32_t cnt = sh.nT;
    while (cnt > 0) {
        ++bfs_levels;
        const uint32_t nl = level + 1;
        for (uint32_t i = tid; i < cnt; i += kMfThreads) {
            uint32_t w = T.get(i);
            if (w < hi) {  // back arc (w+1) -> w is always residual
                if (atomicCAS(&na.d_cur[w + 1], kLabelInf, nl) == kLabelInf)
                    q_append(N, &sh.nN, w + 1);
            }
            if (w > lo && na.g[w] > 0) {  // reverse of back arc w -> w-1
                if (atomicCAS(&na.d_cur[w - 1], kLabelInf, nl) == kLabelInf)
                    q_append(N, &sh.nN, w - 1);
            }
            for (uint32_t k = bg.in_ptr[w], ke = bg.in_ptr[w + 1]; k < ke; ++k) {
                uint32_t b = bg.in_bid[k];
                if (bg.f[b] < bg.b_mult[b]) {
                    uint32_t u = bg.b_s[b];
                    if (atomicCAS(&na.d_cur[u], kLabelInf, nl) == kLabelInf)
                        q_append(N, &sh.nN, u);
                }
            }
            for (uint32_t b = bg.out_ptr[w], be = bg.out_ptr[w + 1]; b < be; ++b) {
                if (bg.f[b] > 0) {
                    uint32_t u = bg.b_t[b];
                    if (atomicCAS(&na.d_cur[u], kLabelInf, nl) == kLabelInf)
                        q_append(N, &sh.nN, u);
                }
            }
        }
        __syncthreads();
        cnt = sh.nN;
        __syncthreads();
        if (tid == 0) {
            sh.nT = cnt;
            sh.nN = 0;
        }
        Queue<QCAP> tmp = T;
        T = N;
        N = tmp;
        ++level;
        __syncthreads();
    }
    for (uint32_t v = lo + tid; v <= hi; v += kMfThreads) na.d_snap[v] = na.d_cur[v];
    if (tid == 0) {
        sh.nT = 0;
        sh.nN = 0;
    }
    __syncthreads();
    return level;
}

template <int THREADS, uint32_t QCAP, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
k_maxflow(NodeArrays na, BundleGraph bg, const uint32_t* __restrict__ comp_lo,
          const uint32_t* __restrict__ comp_hi, uint32_t n_comp, uint32_t* work_counter,
          uint32_t* qF_g, uint32_t* qT_g, uint32_t* qN_g, SolveParams P,
          CompStats* __restrict__ stats) {
    constexpr uint32_t kMfThreads = THREADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MfShared<QCAP>& sh = *reinterpret_cast<MfShared<QCAP>*>(smem_raw);
    const uint32_t tid = threadIdx.x;

    for (;;) {
        if (tid == 0) sh.comp = atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t c = sh.comp;
        if (c >= n_comp) break;
        const uint32_t lo = comp_lo[c], hi = comp_hi[c];
        const uint32_t ncomp = hi - lo + 1;
        Queue<QCAP> F{sh.qa, qF_g + lo}, T{sh.qb, qT_g + lo}, N{sh.qc, qN_g + lo};
        if (tid == 0) {
            sh.nF = 0;
            sh.relabels_since = 0;
            sh.pushes = 0;
            sh.relabels = 0;
            sh.sink_flow = 0;
            sh.stuck = 0;
        }
        unsigned long long bfs_levels = 0, grs = 1, rounds = 0, max_frontier = 0, frontier_sum = 0;
        const long long t_begin = clock64();
        unsigned long long my_pushes = 0, my_relabels = 0;
        long long my_sink = 0, my_stuck = 0;
        uint32_t last_levels = mf_global_relabel<THREADS, QCAP>(na, bg, lo, hi, T, N, sh, bfs_levels);
        for (uint32_t v = lo + tid; v <= hi; v += kMfThreads) {
            if (na.e[v] > 0) {
                na.stamp[v] = 1;
                q_append(F, &sh.nF, v);
            }
        }
        __syncthreads();
        uint32_t round = 0;
        unsigned long long rounds_since = 0;
        for (;;) {
            const uint32_t cntF = sh.nF;
            if (cntF == 0) break;
            if (P.max_rounds && rounds >= P.max_rounds) break;
            unsigned long long interval = (unsigned long long)last_levels * P.gr_levels_pct / 100;
            if (interval < P.gr_interval_min) interval = P.gr_interval_min;
            if (rounds_since >= interval &&
                (unsigned long long)sh.relabels_since * 100 >=
                    (unsigned long long)P.gr_relabel_pct * ncomp) {
                __syncthreads();  // everyone has read relabels_since
                last_levels = mf_global_relabel<THREADS, QCAP>(na, bg, lo, hi, T, N, sh, bfs_levels);
                ++grs;
                if (tid == 0) sh.relabels_since = 0;
                rounds_since = 0;
                __syncthreads();
            }
            ++round;
            ++rounds;
            ++rounds_since;
            if (cntF > max_frontier) max_frontier = cntF;
            frontier_sum += cntF;

            // ---------------- phase A: pushes ----------------
            for (uint32_t i = tid; i < cntF; i += kMfThreads) {
                const uint32_t v = F.get(i);
                const uint32_t dv = na.d_cur[v];
                na.d_snap[v] = dv;  // re-sync the snapshot of a node relabelled last round
                if (dv >= kLabelInf) continue;
                int32_t ex = na.e[v];
                auto give = [&](uint32_t w, int32_t dl) {
                    atomicAdd(&na.eadd[w], dl);
                    if (atomicExch(&na.stamp[w], round) != round) q_append(T, &sh.nT, w);
                    ++my_pushes;
                };
                if (dv == 1) {  // 1. sink arc
                    int32_t s = na.snk[v];
                    if (s > 0) {
                        int32_t dl = min(ex, s);
                        na.snk[v] = s - dl;
                        ex -= dl;
                        my_sink += dl;
                        ++my_pushes;
                    }
                }
                // 2. own bundles, farthest end first
                {
                    const uint32_t ob = bg.out_ptr[v];
                    for (uint32_t b = bg.out_ptr[v + 1]; ex > 0 && b-- > ob;) {
                        const uint32_t t = bg.b_t[b];
                        if (na.d_cur[t] + 1 != dv) continue;
                        const uint32_t fb = bg.f[b];
                        const uint32_t r = bg.b_mult[b] - fb;
                        if (r == 0) continue;
                        const int32_t dl = (int32_t)min((uint32_t)ex, r);
                        bg.f[b] = fb + dl;
                        ex -= dl;
                        give(t, dl);
                    }
                }
                // 3. cancel back-flow towards the right neighbour
                if (ex > 0 && v < hi && na.d_cur[v + 1] + 1 == dv) {
                    const int32_t gr = na.g[v + 1];
                    if (gr > 0) {
                        const int32_t dl = min(ex, gr);
                        na.g[v + 1] = gr - dl;
                        ex -= dl;
                        give(v + 1, dl);
                    }
                }
                // 4. back arc to the left neighbour (infinite capacity)
                if (ex > 0 && v > lo && na.d_cur[v - 1] + 1 == dv) {
                    na.g[v] += ex;
                    give(v - 1, ex);
                    ex = 0;
                }
                // 5. cancel flow on incoming bundles, nearest start first
                if (ex > 0) {
                    const uint32_t ib = bg.in_ptr[v];
                    for (uint32_t k = bg.in_ptr[v + 1]; ex > 0 && k-- > ib;) {
                        const uint32_t b = bg.in_bid[k];
                        const uint32_t s = bg.b_s[b];
                        if (na.d_cur[s] + 1 != dv) continue;
                        const uint32_t fb = bg.f[b];
                        if (fb == 0) continue;
                        const int32_t dl = (int32_t)min((uint32_t)ex, fb);
                        bg.f[b] = fb - dl;
                        ex -= dl;
                        give(s, dl);
                    }
                }
                na.e[v] = ex;
            }
            __syncthreads();
            const uint32_t cntT = sh.nT;

            // ---------------- phase B: merge, relabel from snapshot, next frontier ----------------
            for (uint32_t i = tid; i < cntF + cntT; i += kMfThreads) {
                const bool in_front = i < cntF;
                const uint32_t w = in_front ? F.get(i) : T.get(i - cntF);
                const int32_t left = na.e[w];
                const uint32_t dw = na.d_cur[w];
                bool frozen = dw >= kLabelInf;
                if (in_front && left > 0 && !frozen) {
                    uint32_t mn = kLabelInf;
                    if (na.snk[w] > 0) mn = 0;
                    for (uint32_t b = bg.out_ptr[w], be = bg.out_ptr[w + 1]; b < be; ++b)
                        if (bg.f[b] < bg.b_mult[b]) mn = min(mn, na.d_snap[bg.b_t[b]]);
                    if (w < hi && na.g[w + 1] > 0) mn = min(mn, na.d_snap[w + 1]);
                    if (w > lo) mn = min(mn, na.d_snap[w - 1]);
                    for (uint32_t k = bg.in_ptr[w], ke = bg.in_ptr[w + 1]; k < ke; ++k) {
                        const uint32_t b = bg.in_bid[k];
                        if (bg.f[b] > 0) mn = min(mn, na.d_snap[bg.b_s[b]]);
                    }
                    na.d_cur[w] = mn >= kLabelInf ? kLabelInf : mn + 1;
                    ++my_relabels;
                    atomicAdd(&sh.relabels_since, 1u);
                    frozen = false;  // stays queued one more round so its snapshot is re-synced
                }
                const int32_t add = na.eadd[w];
                const int32_t tot = left + add;
                if (add) na.eadd[w] = 0;
                na.e[w] = tot;
                if (tot > 0) {
                    if (frozen) {
                        my_stuck += tot;
                    } else {
                        na.stamp[w] = round + 1;
                        q_append(N, &sh.nN, w);
                    }
                }
            }
            __syncthreads();
            {
                Queue<QCAP> tmp = F;
                F = N;
                N = tmp;
            }
            const uint32_t nn = sh.nN;
            __syncthreads();
            if (tid == 0) {
                sh.nF = nn;
                sh.nN = 0;
                sh.nT = 0;
            }
            __syncthreads();
        }
        // ---- per-component statistics ----
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            my_pushes += __shfl_xor_sync(0xffffffffu, my_pushes, o);
            my_relabels += __shfl_xor_sync(0xffffffffu, my_relabels, o);
            my_sink += __shfl_xor_sync(0xffffffffu, my_sink, o);
            my_stuck += __shfl_xor_sync(0xffffffffu, my_stuck, o);
        }
        if (lane_id() == 0) {
            atomicAdd(&sh.pushes, my_pushes);
            atomicAdd(&sh.relabels, my_relabels);
            atomicAdd((unsigned long long*)&sh.sink_flow, (unsigned long long)my_sink);
            atomicAdd((unsigned long long*)&sh.stuck, (unsigned long long)my_stuck);
        }
        __syncthreads();
        if (tid == 0) {
            CompStats cs;
            cs.rounds = rounds;
            cs.pushes = sh.pushes;
            cs.relabels = sh.relabels;
            cs.grs = grs;
            cs.bfs_levels = bfs_levels;
            cs.max_frontier = max_frontier;
            cs.sink_flow = sh.sink_flow;
            cs.stuck = sh.stuck;
            cs.cycles = (unsigned long long)(clock64() - t_begin);
            cs.frontier_sum = frontier_sum;
            stats[c] = cs;
        }
        __syncthreads();
    }
}

}  // namespace gds
