// radix_sort.cuh — stable LSD radix sort of (key, uint32 value) pairs, 8 bits per pass.
// Stability is what makes read selection deterministic: after sorting by (start node, length)
// the reads of one bundle appear in ascending read index.
//
// Per pass: k_rs_hist (per-tile digit counts, digit-major) -> exclusive scan -> k_rs_scatter
// (stable in-tile ranks from warp match_any + per-warp running histograms in shared memory).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace gds {

constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 keys per tile

// Key sources: a plain array, or a functor that builds the key from the reads (first pass).
template <typename K>
struct ArrayKeys {
    const K* keys;
    __device__ __forceinline__ void begin_tile(size_t, size_t) const {}
    __device__ __forceinline__ K get(size_t i) const { return keys[i]; }
    __device__ __forceinline__ K get(size_t i, uint32_t& owner) const {
        owner = (uint32_t)i;
        return keys[i];
    }
};

template <typename K, typename KS>
__global__ void __launch_bounds__(kRsThreads)
k_rs_hist(KS ks, size_t n, int shift, uint32_t n_tiles, uint32_t* __restrict__ tile_hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    size_t base = (size_t)blockIdx.x * kRsTile;
    ks.begin_tile(base, n);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
        size_t i = base + (size_t)k * kRsThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(ks.get(i) >> shift) & 255u], 1u);
    }
    __syncthreads();
    tile_hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// vals_in == nullptr means "value = what the key source says" (first pass: the owner read).
template <typename K, typename KS>
__global__ void __launch_bounds__(kRsThreads)
k_rs_scatter(KS ks, const uint32_t* __restrict__ vals_in,
             K* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n, int shift,
             uint32_t n_tiles, const uint32_t* __restrict__ tile_off) {
    __shared__ uint32_t whist[kRsWarps][256];  // running per-warp digit counts
    __shared__ uint32_t gbase[256];            // global offset of this tile's first key per digit
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    for (int i = threadIdx.x; i < kRsWarps * 256; i += kRsThreads) (&whist[0][0])[i] = 0;
    gbase[threadIdx.x] = tile_off[(size_t)threadIdx.x * n_tiles + blockIdx.x];
    ks.begin_tile((size_t)blockIdx.x * kRsTile, n);
    __syncthreads();

    // warp w owns the contiguous chunk [w*32*ITEMS, (w+1)*32*ITEMS) of the tile; step k covers 32
    // consecutive keys, so (warp, step, lane) order == input order.
    const size_t wbase = (size_t)blockIdx.x * kRsTile + (size_t)warp * 32 * kRsItems;
    K key[kRsItems];
    uint32_t rank[kRsItems];
    uint32_t own[kRsItems];
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
        size_t i = wbase + (size_t)k * 32 + lane;
        bool valid = i < n;
        own[k] = 0;
        key[k] = valid ? ks.get(i, own[k]) : (K)0;
        uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
        uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 256u + lane);
        uint32_t before = __popc(peers & lanemask_lt());
        uint32_t prev = 0;
        if (valid) prev = whist[warp][d];
        __syncwarp();
        if (valid && before == 0) whist[warp][d] = prev + __popc(peers);
        __syncwarp();
        rank[k] = prev + before;
    }
    __syncthreads();
    {   // exclusive prefix over warps for digit == threadIdx.x
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            uint32_t c = whist[w][threadIdx.x];
            whist[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
        size_t i = wbase + (size_t)k * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
            size_t dst = (size_t)gbase[d] + whist[warp][d] + rank[k];
            keys_out[dst] = key[k];
            vals_out[dst] = vals_in ? vals_in[i] : own[k];
        }
    }
}

struct RadixTemp {
    DevBuf hist;
    ScanTemp scan;
};

// Sorts n pairs by the low `bits` bits of the key.  Buffers ping-pong; returns 0 if the result
// is in (keys_a, vals_a), 1 if in (keys_b, vals_b).  If iota_first, vals_a is ignored as input and
// the values are the original element indices.
// If first_ks is non-null the first pass takes its keys from *first_ks instead of keys_a.
template <typename K, typename KS0 = ArrayKeys<K>>
inline int radix_sort_pairs(K* keys_a, uint32_t* vals_a, K* keys_b, uint32_t* vals_b, size_t n,
                            int bits, bool iota_first, RadixTemp& tmp, cudaStream_t st,
                            int* passes_out = nullptr, const KS0* first_ks = nullptr) {
    int passes = (bits + 7) / 8;
    if (passes == 0) passes = 1;  // still need vals materialised
    if (passes_out) *passes_out = passes;
    if (n == 0) return 0;
    uint32_t n_tiles = (uint32_t)((n + kRsTile - 1) / kRsTile);
    uint32_t* hist = tmp.hist.get<uint32_t>((size_t)256 * n_tiles);
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        K* kin = cur ? keys_b : keys_a;
        K* kout = cur ? keys_a : keys_b;
        uint32_t* vin = cur ? vals_b : vals_a;
        uint32_t* vout = cur ? vals_a : vals_b;
        int shift = 8 * p;
        const uint32_t* vsrc = (p == 0 && iota_first) ? nullptr : vin;
        if (p == 0 && first_ks) {
            {
                KScope ks("rs_hist_reads", 8ull * n, st);
                k_rs_hist<K, KS0><<<n_tiles, kRsThreads, 0, st>>>(*first_ks, n, shift, n_tiles, hist);
                GDS_KERNEL_CHECK();
            }
            exclusive_scan_u32(hist, hist, (size_t)256 * n_tiles, tmp.scan, st);
            {
                KScope ks("rs_scatter_reads", (8ull + sizeof(K) + 4) * n, st);
                k_rs_scatter<K, KS0><<<n_tiles, kRsThreads, 0, st>>>(*first_ks, vsrc, kout, vout, n,
                                                                     shift, n_tiles, hist);
                GDS_KERNEL_CHECK();
            }
        } else {
            ArrayKeys<K> ak{kin};
            {
                KScope ks("rs_hist", sizeof(K) * (unsigned long long)n, st);
                k_rs_hist<K, ArrayKeys<K>><<<n_tiles, kRsThreads, 0, st>>>(ak, n, shift, n_tiles, hist);
                GDS_KERNEL_CHECK();
            }
            exclusive_scan_u32(hist, hist, (size_t)256 * n_tiles, tmp.scan, st);
            {
                KScope ks("rs_scatter", (2ull * sizeof(K) + (vsrc ? 8 : 4)) * n, st);
                k_rs_scatter<K, ArrayKeys<K>><<<n_tiles, kRsThreads, 0, st>>>(ak, vsrc, kout, vout,
                                                                              n, shift, n_tiles, hist);
                GDS_KERNEL_CHECK();
            }
        }
        cur ^= 1;
    }
    return cur;
}

}  // namespace gds
