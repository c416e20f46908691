// graph.cuh — K2 (coverage flow graph over reference positions) and K4 (zero-coverage cuts).
//
// Network restated from quasi_mcp_cpu_max_flow_solver.cpp:30-87 (identical construction in
// quasi_mcp_cuda_max_flow_solver.cu:157-259), on read BUNDLES (all reads with equal
// (start node, end node) share one arc of capacity = multiplicity):
//   * sort reads by key (start node, length)  -> bundles, ascending read index inside a bundle
//   * difference array over nodes from the bundles (+mult at s, -mult at t) -> scan -> coverage,
//     capped coverage, demand, initial excess (source arcs pre-saturated), sink capacities
//   * out-CSR (bundles by start) and in-CSR (bundle ids by end, ascending start)
//   * components = maximal runs of covered positions (a back arc over an uncovered position can
//     carry no flow in any maximum flow, so it is cut)
#pragma once
#include "common.cuh"
#include "prep.cuh"
#include "scan.cuh"

namespace gds {

__device__ __forceinline__ uint32_t* tile_sample_range() {
    __shared__ uint32_t r[2];
    return r;
}

// First-pass key source of the read sort: key = (node(start) << lenbits) | (len - minlen)
template <typename K>
struct ReadKeys {
    const uint32_t* S;
    const uint32_t* E;
    const uint64_t* off;
    const uint32_t* base;
    uint32_t n_samples;
    int lenbits;
    uint32_t minlen;
    __device__ __forceinline__ void begin_tile(size_t first, size_t n) const {
        if (threadIdx.x == 0) {
            uint32_t* r = tile_sample_range();
            size_t last = min(first + (size_t)blockDim.x * 64, n) - 1;  // >= any index of the tile
            if (n_samples == 1) {
                r[0] = r[1] = 0;
            } else {
                r[0] = find_sample(off, n_samples, first);
                r[1] = find_sample(off, n_samples, last);
            }
        }
    }
    __device__ __forceinline__ K get(size_t i) const {
        const uint32_t* r = tile_sample_range();
        uint32_t k = r[0];
        if (r[0] != r[1]) k = find_sample(off, n_samples, i);
        uint32_t s = S[i], e = E[i];
        return ((K)(base[k] + s) << lenbits) | (K)(e - s + 1 - minlen);
    }
};

constexpr int kHeadThreads = 1024;
constexpr int kHeadItems = 4;
constexpr int kHeadTile = kHeadThreads * kHeadItems;

template <typename K>
__global__ void __launch_bounds__(kHeadThreads)
k_heads_count(const K* __restrict__ keys, size_t n, uint32_t* __restrict__ tile_counts) {
    size_t base = (size_t)blockIdx.x * kHeadTile + (size_t)threadIdx.x * kHeadItems;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < kHeadItems; ++k) {
        size_t j = base + k;
        if (j < n) c += (j == 0 || keys[j] != keys[j - 1]) ? 1u : 0u;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ uint32_t tot;
    if (threadIdx.x == 0) tot = 0;
    __syncthreads();
    if (lane_id() == 0 && c) atomicAdd(&tot, c);
    __syncthreads();
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = tot;
}

template <typename K>
__global__ void __launch_bounds__(kHeadThreads)
k_heads_write(const K* __restrict__ keys, size_t n, const uint32_t* __restrict__ tile_offs,
              uint32_t* __restrict__ b_first, K* __restrict__ b_key) {
    __shared__ uint32_t total;
    size_t base = (size_t)blockIdx.x * kHeadTile + (size_t)threadIdx.x * kHeadItems;
    bool head[kHeadItems];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < kHeadItems; ++k) {
        size_t j = base + k;
        head[k] = j < n && (j == 0 || keys[j] != keys[j - 1]);
        c += head[k] ? 1u : 0u;
    }
    uint32_t ex = block_excl_scan(c, &total) + tile_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kHeadItems; ++k) {
        if (head[k]) {
            b_first[ex] = (uint32_t)(base + k);
            b_key[ex] = keys[base + k];
            ++ex;
        }
    }
}

// Per bundle: decode (s, t), multiplicity, and accumulate the node-level difference array and the
// CSR degree counters.  One atomic triple per BUNDLE, not per read.
template <typename K>
__global__ void __launch_bounds__(256)
k_bundle_fill(const K* __restrict__ b_key, uint32_t* __restrict__ b_first, uint32_t B, uint32_t N,
              int lenbits, uint32_t minlen, uint32_t* __restrict__ b_s, uint32_t* __restrict__ b_t,
              uint32_t* __restrict__ b_mult, int32_t* __restrict__ diff,
              uint32_t* __restrict__ outdeg, uint32_t* __restrict__ indeg) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    K key = b_key[b];
    uint32_t s = (uint32_t)(key >> lenbits);
    uint32_t len = (uint32_t)(key & (((K)1 << lenbits) - 1)) + minlen;
    uint32_t t = s + len;
    uint32_t nxt = (b + 1 < B) ? b_first[b + 1] : N;
    uint32_t mult = nxt - b_first[b];
    b_s[b] = s;
    b_t[b] = t;
    b_mult[b] = mult;
    atomicAdd(&diff[s], (int32_t)mult);
    atomicAdd(&diff[t], -(int32_t)mult);
    atomicAdd(&outdeg[s], 1u);
    atomicAdd(&indeg[t], 1u);
}

struct NodeArrays {
    uint32_t* d_cur;   // labels
    uint32_t* d_snap;  // label snapshot read by relabels
    int32_t* e;        // excess
    int32_t* eadd;     // excess received during the current round
    int32_t* snk;      // remaining sink-arc capacity
    int32_t* g;        // flow on the back arc v -> v-1
    uint32_t* stamp;   // round in which v was queued
};

// covL[v] = coverage of the position left of node v = exclusive prefix; covR = inclusive prefix.
// totals[0] (u64) accumulates F*.
__global__ void __launch_bounds__(256)
k_node_finalize(const uint32_t* __restrict__ excl, const int32_t* __restrict__ diff,
                uint32_t n_nodes, uint32_t M, NodeArrays na, uint32_t* __restrict__ comp_start,
                uint32_t* __restrict__ comp_end, uint32_t* __restrict__ cov_capped_out,
                int32_t* __restrict__ demand_out, unsigned long long* __restrict__ totals) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long src = 0;
    if (v < n_nodes) {
        uint32_t covL = excl[v];
        uint32_t covR = covL + (uint32_t)diff[v];
        int32_t dem = (int32_t)min(covL, M) - (int32_t)min(covR, M);
        na.e[v] = dem < 0 ? -dem : 0;
        na.snk[v] = dem > 0 ? dem : 0;
        na.g[v] = 0;
        na.eadd[v] = 0;
        na.stamp[v] = 0;
        na.d_cur[v] = kLabelInf;
        na.d_snap[v] = kLabelInf;
        comp_start[v] = (covL == 0 && covR > 0) ? 1u : 0u;
        comp_end[v] = (covL > 0 && covR == 0) ? 1u : 0u;
        if (cov_capped_out) cov_capped_out[v] = min(covR, M);
        if (demand_out) demand_out[v] = dem;
        if (dem < 0) src = (unsigned long long)(-dem);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) src += __shfl_xor_sync(0xffffffffu, src, o);
    if (lane_id() == 0 && src) atomicAdd(&totals[0], src);
}

__global__ void __launch_bounds__(256)
k_comp_write(const uint32_t* __restrict__ comp_start, const uint32_t* __restrict__ comp_end,
             const uint32_t* __restrict__ start_idx, const uint32_t* __restrict__ end_idx,
             uint32_t n_nodes, uint32_t* __restrict__ comp_lo, uint32_t* __restrict__ comp_hi) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    if (comp_start[v]) comp_lo[start_idx[v]] = v;
    if (comp_end[v]) comp_hi[end_idx[v]] = v;
}

}  // namespace gds
