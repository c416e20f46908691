// radix_sort.cuh — stable, SEGMENTED LSD radix sort of (key, uint32 value) pairs, 8 bits per pass.
//
// Stability is what makes read selection deterministic: after sorting by (start node, length)
// the reads of one bundle appear in ascending read index.
//
// Segmented: the items come pre-grouped (group = sample; its reads are contiguous in the input),
// so the group id is an implicit, already sorted key prefix and only the key bits INSIDE a group
// are sorted.  Tiles of kRsTile items never straddle a group; the per-tile digit histograms are
// laid out [group][digit][tile in group], so one exclusive scan over that array yields, for every
// (group, digit, tile), the global position of the tile's first key with that digit.  A batch of
// 512 samples x 30 kb therefore sorts 15-bit keys in 2 passes instead of 24-bit keys in 3.
//
// Per pass: k_rs_hist (per-tile digit counts) -> exclusive scan -> k_rs_scatter.  The scatter
// ranks keys with warp match_any + per-warp running counts (stable), reorders the whole tile by
// digit in shared memory and writes each digit run as one contiguous, coalesced segment
// (a direct scatter wrote 30 sectors per 32-lane store; this writes ~5).
#pragma once
#include "common.cuh"
#include "prep.cuh"
#include "scan.cuh"

namespace gds {

constexpr int kRsThreads = 512;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 8192 keys per tile

// Where the tiles are.  n_groups == 1: items [0, n_items) in tiles of kRsTile.  Otherwise group g
// owns items [item_off[g], item_off[g+1]) and tiles [tile_off[g], tile_off[g+1]).
struct TileMap {
    const uint32_t* tile_off;  // [n_groups+1], device
    const uint64_t* item_off;  // [n_groups+1], device
    uint32_t n_groups;
    uint32_t n_tiles;
    size_t n_items;
};

struct TilePos {
    size_t first;        // first item of the tile
    uint32_t n_valid;    // items in the tile
    uint32_t hist_base;  // index of hist[group][digit 0][tile 0]
    uint32_t tiles_g;    // tiles in the group
    uint32_t tile_in_g;
};

// executed by every thread of the block (uniform result); the binary search is ~9 L1-cached steps
__device__ __forceinline__ TilePos locate_tile(const TileMap& tm, uint32_t tile) {
    TilePos p;
    if (tm.n_groups == 1) {
        p.first = (size_t)tile * kRsTile;
        p.n_valid = (uint32_t)min((size_t)kRsTile, tm.n_items - p.first);
        p.hist_base = 0;
        p.tiles_g = tm.n_tiles;
        p.tile_in_g = tile;
        return p;
    }
    uint32_t lo = 0, hi = tm.n_groups;  // largest g with tile_off[g] <= tile
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (tm.tile_off[mid] <= tile) lo = mid;
        else hi = mid;
    }
    const uint32_t t0 = tm.tile_off[lo];
    const uint64_t i0 = tm.item_off[lo], i1 = tm.item_off[lo + 1];
    p.tile_in_g = tile - t0;
    p.tiles_g = tm.tile_off[lo + 1] - t0;
    p.first = (size_t)i0 + (size_t)p.tile_in_g * kRsTile;
    p.n_valid = (uint32_t)min((uint64_t)kRsTile, i1 - p.first);
    p.hist_base = 256u * t0;
    return p;
}

// Key sources: a plain array, or a functor that builds the key from the reads (first pass).
template <typename K>
struct ArrayKeys {
    const K* keys;
    __device__ __forceinline__ K get(size_t i) const { return keys[i]; }
    __device__ __forceinline__ uint32_t owner(size_t i) const { return (uint32_t)i; }
};

template <typename K, typename KS>
__global__ void __launch_bounds__(kRsThreads)
k_rs_hist(KS ks, TileMap tm, int shift, uint32_t* __restrict__ tile_hist) {
    __shared__ uint32_t h[kRsWarps][256];  // per-warp counts: conflicts only inside a warp
    for (int i = threadIdx.x; i < kRsWarps * 256; i += kRsThreads) (&h[0][0])[i] = 0;
    const TilePos tp = locate_tile(tm, blockIdx.x);
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5;
    K key[kRsItems];
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
        uint32_t j = (uint32_t)k * kRsThreads + threadIdx.x;
        key[k] = j < tp.n_valid ? ks.get(tp.first + j) : (K)0;
    }
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
        uint32_t j = (uint32_t)k * kRsThreads + threadIdx.x;
        if (j < tp.n_valid) atomicAdd(&h[warp][(uint32_t)(key[k] >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) s += h[w][threadIdx.x];
        tile_hist[(size_t)tp.hist_base + (size_t)threadIdx.x * tp.tiles_g + tp.tile_in_g] = s;
    }
}

template <typename K>
struct RsSmem {
    K skey[kRsTile];
    uint32_t sval[kRsTile];
    uint16_t whist[kRsWarps][256];  // running per-warp digit counts (<= 512 per warp)
    uint32_t dstart[256];           // first slot of each digit in the reordered tile
    uint32_t gbase[256];            // global position of slot 0 of each digit's run, minus dstart
    uint32_t wsum[8];
};

// vals_in == nullptr means "value = what the key source says" (first pass: the owner read).
template <typename K, typename KS, int MIN_CTAS>
__global__ void __launch_bounds__(kRsThreads, MIN_CTAS)
k_rs_scatter(KS ks, const uint32_t* __restrict__ vals_in, K* __restrict__ keys_out,
             uint32_t* __restrict__ vals_out, TileMap tm, int shift,
             const uint32_t* __restrict__ tile_off_scanned) {
    extern __shared__ __align__(16) unsigned char rs_smem_raw[];
    RsSmem<K>& sm = *reinterpret_cast<RsSmem<K>*>(rs_smem_raw);
    const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
    for (int i = threadIdx.x; i < kRsWarps * 256 / 2; i += kRsThreads)
        reinterpret_cast<uint32_t*>(&sm.whist[0][0])[i] = 0;
    const TilePos tp = locate_tile(tm, blockIdx.x);
    if (threadIdx.x < 256)
        sm.gbase[threadIdx.x] = tile_off_scanned[(size_t)tp.hist_base +
                                                 (size_t)threadIdx.x * tp.tiles_g + tp.tile_in_g];

    // warp w owns the contiguous chunk [w*32*ITEMS, (w+1)*32*ITEMS) of the tile; row k covers 32
    // consecutive keys, so (warp, row, lane) order == input order.  All loads are issued first.
    const uint32_t wofs = warp * 32 * kRsItems;
    K key[kRsItems];
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
        uint32_t j = wofs + (uint32_t)k * 32 + lane;
        key[k] = j < tp.n_valid ? ks.get(tp.first + j) : (K)0;
    }
    __syncthreads();  // whist zeroed
    uint32_t rank2[kRsItems / 2];  // two 16-bit ranks per register
#pragma unroll
    for (int k = 0; k < kRsItems / 2; ++k) rank2[k] = 0;
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
        uint32_t j = wofs + (uint32_t)k * 32 + lane;
        bool valid = j < tp.n_valid;
        uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
        uint32_t peers = __match_any_sync(0xffffffffu, valid ? d : 256u + lane);
        uint32_t before = __popc(peers & lanemask_lt());
        uint32_t prev = valid ? sm.whist[warp][d] : 0;
        __syncwarp();
        if (valid && before == 0) sm.whist[warp][d] = (uint16_t)(prev + __popc(peers));
        __syncwarp();
        rank2[k >> 1] |= (prev + before) << (16 * (k & 1));
    }
    __syncthreads();
    uint32_t tot = 0;
    if (threadIdx.x < 256) {  // exclusive prefix over warps for digit == threadIdx.x
#pragma unroll
        for (int w = 0; w < kRsWarps; ++w) {
            uint32_t c = sm.whist[w][threadIdx.x];
            sm.whist[w][threadIdx.x] = (uint16_t)tot;
            tot += c;
        }
        // exclusive scan of the 256 digit totals (8 warps)
        uint32_t incl = warp_incl_scan(tot);
        if (lane == 31) sm.wsum[warp] = incl;
        sm.dstart[threadIdx.x] = incl - tot;  // warp-local for now
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        uint32_t add = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) add += (w < (int)warp) ? sm.wsum[w] : 0u;
        uint32_t ds = sm.dstart[threadIdx.x] + add;
        sm.dstart[threadIdx.x] = ds;
        sm.gbase[threadIdx.x] -= ds;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kRsItems; ++k) {
        uint32_t j = wofs + (uint32_t)k * 32 + lane;
        if (j < tp.n_valid) {
            uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
            uint32_t p = sm.dstart[d] + sm.whist[warp][d] + ((rank2[k >> 1] >> (16 * (k & 1))) & 0xffffu);
            sm.skey[p] = key[k];
            sm.sval[p] = vals_in ? vals_in[tp.first + j] : ks.owner(tp.first + j);
        }
    }
    __syncthreads();
#pragma unroll 4
    for (uint32_t j = threadIdx.x; j < tp.n_valid; j += kRsThreads) {
        K kk = sm.skey[j];
        uint32_t d = (uint32_t)(kk >> shift) & 255u;
        size_t dst = (size_t)(sm.gbase[d] + j);
        keys_out[dst] = kk;
        vals_out[dst] = sm.sval[j];
    }
}

struct RadixTemp {
    DevBuf hist;
    ScanTemp scan;
    bool attr32 = false, attr64 = false;
};

inline uint32_t tiles_for(size_t n) { return (uint32_t)((n + kRsTile - 1) / kRsTile); }

// Sorts the pairs by the low `bits` bits of the key inside every group of `tm`.  Buffers
// ping-pong; returns 0 if the result is in (keys_a, vals_a), 1 if in (keys_b, vals_b).
// The first pass takes key and value from *first_ks when it is non-null, else from keys_a with
// value = item index.
template <typename K, typename KS0 = ArrayKeys<K>>
inline int radix_sort_pairs(K* keys_a, uint32_t* vals_a, K* keys_b, uint32_t* vals_b,
                            const TileMap& tm, int bits, RadixTemp& tmp, cudaStream_t st,
                            int* passes_out = nullptr, const KS0* first_ks = nullptr) {
    int passes = (bits + 7) / 8;
    if (passes == 0) passes = 1;  // still need vals materialised
    if (passes_out) *passes_out = passes;
    const size_t n = tm.n_items;
    if (n == 0 || tm.n_tiles == 0) return 0;
    const uint32_t n_tiles = tm.n_tiles;
    uint32_t* hist = tmp.hist.get<uint32_t>((size_t)256 * n_tiles);
    constexpr int kMinCtas = sizeof(K) == 4 ? 2 : 1;
    constexpr int smem = (int)sizeof(RsSmem<K>);
    bool& attr = sizeof(K) == 4 ? tmp.attr32 : tmp.attr64;
    if (!attr) {
        GDS_CUDA(cudaFuncSetAttribute(k_rs_scatter<K, KS0, kMinCtas>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GDS_CUDA(cudaFuncSetAttribute(k_rs_scatter<K, ArrayKeys<K>, kMinCtas>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        K* kin = cur ? keys_b : keys_a;
        K* kout = cur ? keys_a : keys_b;
        uint32_t* vin = cur ? vals_b : vals_a;
        uint32_t* vout = cur ? vals_a : vals_b;
        int shift = 8 * p;
        if (p == 0 && first_ks) {
            {
                KScope ks("rs_hist_reads", 8ull * n, st);
                k_rs_hist<K, KS0><<<n_tiles, kRsThreads, 0, st>>>(*first_ks, tm, shift, hist);
                GDS_KERNEL_CHECK();
            }
            exclusive_scan_u32(hist, hist, (size_t)256 * n_tiles, tmp.scan, st);
            {
                KScope ks("rs_scatter_reads", (8ull + sizeof(K) + 4) * n, st);
                k_rs_scatter<K, KS0, kMinCtas><<<n_tiles, kRsThreads, smem, st>>>(
                    *first_ks, nullptr, kout, vout, tm, shift, hist);
                GDS_KERNEL_CHECK();
            }
        } else {
            ArrayKeys<K> ak{kin};
            const uint32_t* vsrc = p == 0 ? nullptr : vin;
            {
                KScope ks("rs_hist", sizeof(K) * (unsigned long long)n, st);
                k_rs_hist<K, ArrayKeys<K>><<<n_tiles, kRsThreads, 0, st>>>(ak, tm, shift, hist);
                GDS_KERNEL_CHECK();
            }
            exclusive_scan_u32(hist, hist, (size_t)256 * n_tiles, tmp.scan, st);
            {
                KScope ks("rs_scatter", (2ull * sizeof(K) + (vsrc ? 8 : 4)) * n, st);
                k_rs_scatter<K, ArrayKeys<K>, kMinCtas><<<n_tiles, kRsThreads, smem, st>>>(
                    ak, vsrc, kout, vout, tm, shift, hist);
                GDS_KERNEL_CHECK();
            }
        }
        cur ^= 1;
    }
    return cur;
}

}  // namespace gds
