// scan.cuh — device-wide exclusive prefix sum over uint32 (wrap-around add, so it also serves
// int32 difference arrays).  Hierarchical reduce-then-scan: deterministic, no atomics.
#pragma once
#include "common.cuh"

namespace gds {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if ((int)lane_id() >= o) v += t;
    }
    return v;
}

// exclusive scan of a block-wide value; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[32];
    uint32_t incl = warp_incl_scan(v);
    uint32_t w = threadIdx.x >> 5;
    if (lane_id() == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        uint32_t nw = (blockDim.x + 31) >> 5;
        uint32_t s = lane_id() < nw ? warp_sums[lane_id()] : 0;
        uint32_t si = warp_incl_scan(s);
        warp_sums[lane_id()] = si - s;  // exclusive warp offsets
        if (lane_id() == 31) *total = si;
    }
    __syncthreads();
    uint32_t r = warp_sums[w] + incl - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_tiles(const uint32_t* in, uint32_t* out /* may alias in: in-place scans */,
             uint32_t* __restrict__ tile_sums, size_t n) {
    __shared__ uint32_t total;
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = base + k < n ? in[base + k] : 0u;
        s += v[k];
    }
    uint32_t ex = block_excl_scan(s, &total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0 && tile_sums) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_add(uint32_t* __restrict__ out, const uint32_t* __restrict__ tile_offs, size_t n) {
    uint32_t off = tile_offs[blockIdx.x];
    size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) out[base + k] += off;
}

struct ScanTemp {
    DevBuf l1, l2;
};

// out may alias in.  total (optional, device pointer) receives nothing here: callers that need
// the grand total read out[n-1] + in[n-1] themselves or scan n+1 elements with a trailing zero.
inline void exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, ScanTemp& tmp,
                               cudaStream_t st) {
    if (n == 0) return;
    KScope ks("scan_u32", 8ull * n, st);  // the whole hierarchical scan counts as one unit
    size_t t1 = (n + kScanTile - 1) / kScanTile;
    if (t1 == 1) {
        k_scan_tiles<<<1, kScanThreads, 0, st>>>(in, out, nullptr, n);
        GDS_KERNEL_CHECK();
        return;
    }
    uint32_t* s1 = tmp.l1.get<uint32_t>(t1);
    k_scan_tiles<<<(unsigned)t1, kScanThreads, 0, st>>>(in, out, s1, n);
    GDS_KERNEL_CHECK();
    size_t t2 = (t1 + kScanTile - 1) / kScanTile;
    if (t2 == 1) {
        k_scan_tiles<<<1, kScanThreads, 0, st>>>(s1, s1, nullptr, t1);
        GDS_KERNEL_CHECK();
    } else {
        uint32_t* s2 = tmp.l2.get<uint32_t>(t2);
        k_scan_tiles<<<(unsigned)t2, kScanThreads, 0, st>>>(s1, s1, s2, t1);
        GDS_KERNEL_CHECK();
        // t2 <= 4096 for n <= 2^36
        k_scan_tiles<<<1, kScanThreads, 0, st>>>(s2, s2, nullptr, t2);
        GDS_KERNEL_CHECK();
        k_scan_add<<<(unsigned)t2, kScanThreads, 0, st>>>(s1, s2, t1);
        GDS_KERNEL_CHECK();
    }
    k_scan_add<<<(unsigned)t1, kScanThreads, 0, st>>>(out, s1, n);
    GDS_KERNEL_CHECK();
}

}  // namespace gds
