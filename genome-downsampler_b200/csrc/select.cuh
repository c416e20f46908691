// select.cuh — K5 (read selection from bundle flows as a bitmap), K6 (coverage verification of
// the kept set by difference array + scan) and the find_pairs bitmap step.
//
// K5 restates obtain_sequence (quasi_mcp_cpu_max_flow_solver.cpp:89-100: keep read i iff
// Flow(i) > 0) and the residual==0 export of quasi_mcp_cuda_max_flow_solver.cu:421-432 on
// bundles: a bundle with flow f keeps its f lowest-index reads.  The sorted
// read order makes those the first f entries of the bundle's slice.
#pragma once
#include "common.cuh"
#include "graph.cuh"

namespace gds {

// one warp per 32 bundles; lane b handles bundle b's slice cooperatively when f is large
__global__ void __launch_bounds__(256)
k_select(const uint32_t* __restrict__ b_first, const BundleRec* __restrict__ bund,
         const uint32_t* __restrict__ sorted_idx, uint32_t B, uint32_t* __restrict__ bitmap,
         unsigned long long* __restrict__ totals /* [1] += n_kept */) {
    const uint32_t b0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u;
    const uint32_t b = b0 + lane_id();
    uint32_t fb = 0, first = 0;
    if (b < B) {
        fb = bund[b].f;
        first = b_first[b];
    }
    // a read cut in two (segment split) can be chosen by both parts: count bits newly set
    unsigned long long kept = 0;
    // small flows: each lane walks its own bundle; large flows: the warp shares the work
    const uint32_t kWide = 64;
    if (fb <= kWide) {
        for (uint32_t r = 0; r < fb; ++r) {
            uint32_t i = sorted_idx[first + r];
            uint32_t bit = 1u << (i & 31);
            kept += (atomicOr(&bitmap[i >> 5], bit) & bit) ? 0 : 1;
        }
    }
    uint32_t wide = __ballot_sync(0xffffffffu, fb > kWide);
    while (wide) {
        int src = __ffs(wide) - 1;
        wide &= wide - 1;
        uint32_t wf = __shfl_sync(0xffffffffu, fb, src);
        uint32_t wfirst = __shfl_sync(0xffffffffu, first, src);
        for (uint32_t r = lane_id(); r < wf; r += 32) {
            uint32_t i = sorted_idx[wfirst + r];
            uint32_t bit = 1u << (i & 31);
            kept += (atomicOr(&bitmap[i >> 5], bit) & bit) ? 0 : 1;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kept += __shfl_xor_sync(0xffffffffu, kept, o);
    if (lane_id() == 0 && kept) atomicAdd(&totals[1], kept);
}

// find_pairs (bam_api.cpp:239-273): mates are adjacent, so "add the mate of every kept read" is a
// per-word bit trick.  n is even per sample, so pairs never straddle words.
__global__ void __launch_bounds__(256)
k_find_pairs(uint32_t* __restrict__ bitmap, size_t n_words) {
    size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t x = bitmap[w];
    uint32_t any = (x & 0x55555555u) | ((x & 0xaaaaaaaau) >> 1);
    bitmap[w] = any | (any << 1);
}

// K6: difference array of the kept reads (bitmap-driven, so it checks the final answer, not the
// flows).  diff must be zeroed.  Reads only the coordinates of kept reads.
__global__ void __launch_bounds__(256)
k_verify_accumulate(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ S,
                    const uint32_t* __restrict__ E, size_t n, const uint64_t* __restrict__ off,
                    const uint32_t* __restrict__ base, uint32_t n_samples,
                    int32_t* __restrict__ diff) {
    size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t n_words = (n + 31) / 32;
    if (w >= n_words) return;
    uint32_t x = bitmap[w];
    while (x) {
        int bit = __ffs(x) - 1;
        x &= x - 1;
        size_t i = w * 32 + bit;
        if (i >= n) break;
        uint32_t k = n_samples == 1 ? 0 : find_sample(off, n_samples, i);
        uint32_t s = base[k] + S[i], t = base[k] + E[i] + 1;
        atomicAdd(&diff[s], 1);
        atomicAdd(&diff[t], -1);
    }
}

// compares min(cov_out, M) with min(cov_in, M) per node; totals[2] += violations
__global__ void __launch_bounds__(256)
k_verify_compare(const uint32_t* __restrict__ excl_out, const int32_t* __restrict__ diff_out,
                 const uint32_t* __restrict__ excl_in, const int32_t* __restrict__ diff_in,
                 uint32_t n_nodes, uint32_t M, unsigned long long* __restrict__ totals) {
    uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bad = 0;
    if (v < n_nodes) {
        uint32_t co = excl_out[v] + (uint32_t)diff_out[v];
        uint32_t ci = excl_in[v] + (uint32_t)diff_in[v];
        bad = min(co, M) != min(ci, M);
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if (lane_id() == 0 && bad) atomicAdd(&totals[2], (unsigned long long)bad);
}

}  // namespace gds
