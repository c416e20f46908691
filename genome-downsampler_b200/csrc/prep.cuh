// prep.cuh — K1: pair filter (predicate + stable compaction) and read validation.
// Restates, for the device, BamApi::should_be_filtered_out (bam_api.cpp:311-332):
//   drop pair unless both mates have quality >= min_mapq, seq_length >= min_len and (FILTER mode)
//   some amplicon [a0,a1] contains both mates (amplicon.cpp:5-7, amplicon_set.cpp:5-9).
// "Some amplicon contains both" == max{a1 : a0 <= min(s1,s2)} >= max(e1,e2); the host ships the
// amplicons sorted by a0 with a running max of a1, so the device does one binary search per pair.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace gds {

// sample lookup: largest k with off[k] <= i   (off[0] = 0, off[n_samples] = total)
__device__ __forceinline__ uint32_t find_sample(const uint64_t* __restrict__ off,
                                                uint32_t n_samples, uint64_t i) {
    uint32_t lo = 0, hi = n_samples;  // invariant: off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}

// Compact transport (gds_reads.start16 / end == NULL): widen to the 32-bit start/end columns the
// pipeline works on.  8 reads per thread, 16-byte accesses.
__global__ void __launch_bounds__(256)
k_expand_reads(const uint16_t* __restrict__ s16, const uint32_t* __restrict__ s32,
               const uint32_t* __restrict__ e32, uint32_t fixed_len, size_t n,
               uint32_t* __restrict__ S, uint32_t* __restrict__ E) {
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i0 >= n) return;
    uint32_t s[8];
    if (i0 + 8 <= n && s16 && (reinterpret_cast<uintptr_t>(s16) & 15) == 0) {
        const uint4 v = ld_stream4(reinterpret_cast<const uint4*>(s16 + i0));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            s[2 * q] = w[q] & 0xffffu;
            s[2 * q + 1] = w[q] >> 16;
        }
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q)
            s[q] = i0 + q < n ? (s16 ? (uint32_t)s16[i0 + q] : s32[i0 + q]) : 0u;
    }
    if (i0 + 8 <= n) {
        if (S != s32) {
            reinterpret_cast<uint4*>(S + i0)[0] = make_uint4(s[0], s[1], s[2], s[3]);
            reinterpret_cast<uint4*>(S + i0)[1] = make_uint4(s[4], s[5], s[6], s[7]);
        }
        if (!e32) {
            const uint32_t d = fixed_len - 1;
            reinterpret_cast<uint4*>(E + i0)[0] = make_uint4(s[0] + d, s[1] + d, s[2] + d, s[3] + d);
            reinterpret_cast<uint4*>(E + i0)[1] = make_uint4(s[4] + d, s[5] + d, s[6] + d, s[7] + d);
        }
    } else {
        for (int q = 0; q < 8 && i0 + q < n; ++q) {
            if (S != s32) S[i0 + q] = s[q];
            if (!e32) E[i0 + q] = s[q] + fixed_len - 1;
        }
    }
}

struct FilterArgs {
    uint32_t min_len, min_mapq;
    uint32_t n_amp;
    const uint32_t* amp_start_sorted;  // ascending
    const uint32_t* amp_end_runmax;    // running max of the matching ends
};

__global__ void __launch_bounds__(256)
k_filter_flags(const uint32_t* __restrict__ start, const uint32_t* __restrict__ end,
               const uint8_t* __restrict__ mapq, const uint32_t* __restrict__ seq_len,
               size_t n_pairs, FilterArgs fa, uint8_t* __restrict__ pair_pass,
               uint32_t* __restrict__ flag32) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs;
         p += (size_t)gridDim.x * blockDim.x) {
        // mates are adjacent: one 8-byte load per array
        uint2 s = reinterpret_cast<const uint2*>(start)[p];
        uint2 e = reinterpret_cast<const uint2*>(end)[p];
        uint2 l = reinterpret_cast<const uint2*>(seq_len)[p];
        uchar2 q = reinterpret_cast<const uchar2*>(mapq)[p];
        bool ok = q.x >= fa.min_mapq && q.y >= fa.min_mapq && l.x >= fa.min_len &&
                  l.y >= fa.min_len;
        if (ok && fa.n_amp) {
            uint32_t lo_pos = min(s.x, s.y), hi_pos = max(e.x, e.y);
            // last amplicon with start <= lo_pos
            int lo = -1, hi = (int)fa.n_amp;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (fa.amp_start_sorted[mid] <= lo_pos) lo = mid;
                else hi = mid;
            }
            ok = lo >= 0 && fa.amp_end_runmax[lo] >= hi_pos;
        }
        pair_pass[p] = ok ? 1 : 0;
        flag32[p] = ok ? 1u : 0u;
    }
}

// pass_idx = exclusive scan of flag32.  Stable compaction: surviving pair p goes to 2*pass_idx[p].
__global__ void __launch_bounds__(256)
k_filter_compact(const uint32_t* __restrict__ start, const uint32_t* __restrict__ end,
                 const uint8_t* __restrict__ pair_pass, const uint32_t* __restrict__ pass_idx,
                 size_t n_pairs, uint32_t* __restrict__ out_start, uint32_t* __restrict__ out_end) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs;
         p += (size_t)gridDim.x * blockDim.x) {
        if (!pair_pass[p]) continue;
        size_t o = pass_idx[p];
        reinterpret_cast<uint2*>(out_start)[o] = reinterpret_cast<const uint2*>(start)[p];
        reinterpret_cast<uint2*>(out_end)[o] = reinterpret_cast<const uint2*>(end)[p];
    }
}

// post-filter read offsets per sample: foff[k] = 2 * (#surviving pairs before pair off[k]/2)
__global__ void k_filter_offsets(const uint64_t* __restrict__ off, uint32_t n_samples,
                                 const uint32_t* __restrict__ pass_idx,
                                 const uint8_t* __restrict__ pair_pass, size_t n_pairs,
                                 uint64_t* __restrict__ foff) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > n_samples) return;
    size_t p = off[k] / 2;
    uint64_t cnt;
    if (n_pairs == 0) cnt = 0;
    else if (p < n_pairs) cnt = pass_idx[p];
    else cnt = (uint64_t)pass_idx[n_pairs - 1] + pair_pass[n_pairs - 1];
    foff[k] = 2 * cnt;
}

// A read [s, e] is acceptable when it lies inside the reference, or when it consumes no reference at
// all (e == s - 1, CIGAR '*': an unmapped mate placed at its mate's position).  The reference keeps
// such a read (Read::Read gives it end = pos + rlen - 1, read.cpp:5-14): its arc start -> end + 1 is a
// self-loop that never carries flow, so it is never kept; here it becomes a bundle of length 0.
__host__ __device__ __forceinline__ bool read_in_range(uint32_t s, uint32_t e, uint32_t L) {
    return (e + 1u == s) ? s < L : (s <= e && e < L);
}

// Validation + read-length range over the (post-filter) reads.  stats[0]=min len, [1]=max len,
// [2]=error count.  One block covers a contiguous chunk of reads; the sample of a read is looked
// up once per block when the chunk lies inside one sample (the common case).
constexpr int kValThreads = 256;
constexpr int kValItems = 16;
constexpr int kValTile = kValThreads * kValItems;

__global__ void __launch_bounds__(kValThreads)
k_validate(const uint32_t* __restrict__ S, const uint32_t* __restrict__ E, size_t n,
           const uint64_t* __restrict__ off, const uint32_t* __restrict__ ref_len,
           uint32_t n_samples, uint32_t* __restrict__ stats) {
    __shared__ uint32_t krange[2];
    const size_t base = (size_t)blockIdx.x * kValTile;
    if (threadIdx.x == 0) {
        size_t last = min(base + kValTile, n) - 1;
        krange[0] = n_samples == 1 ? 0 : find_sample(off, n_samples, base);
        krange[1] = n_samples == 1 ? 0 : find_sample(off, n_samples, last);
    }
    __syncthreads();
    const uint32_t k0 = krange[0];
    const bool one = k0 == krange[1];
    const uint32_t L0 = ref_len[k0];
    uint32_t s[kValItems], e[kValItems];
#pragma unroll
    for (int q = 0; q < kValItems; ++q) {
        size_t i = base + (size_t)q * kValThreads + threadIdx.x;
        size_t ii = i < n ? i : base;
        s[q] = ld_stream(S + ii);
        e[q] = ld_stream(E + ii);
    }
    uint32_t mn = 0xffffffffu, mx = 0, bad = 0;
#pragma unroll
    for (int q = 0; q < kValItems; ++q) {
        size_t i = base + (size_t)q * kValThreads + threadIdx.x;
        if (i >= n) continue;
        uint32_t L = one ? L0 : ref_len[find_sample(off, n_samples, i)];
        if (!read_in_range(s[q], e[q], L)) {
            ++bad;
            continue;
        }
        uint32_t len = e[q] - s[q] + 1;
        mn = min(mn, len);
        mx = max(mx, len);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    __shared__ uint32_t red[3];
    if (threadIdx.x == 0) {
        red[0] = 0xffffffffu;
        red[1] = 0;
        red[2] = 0;
    }
    __syncthreads();
    if (lane_id() == 0) {
        atomicMin(&red[0], mn);
        atomicMax(&red[1], mx);
        if (bad) atomicAdd(&red[2], bad);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (red[0] != 0xffffffffu) atomicMin(&stats[0], red[0]);
        if (red[1]) atomicMax(&stats[1], red[1]);
        if (red[2]) atomicAdd(&stats[2], red[2]);
    }
}

}  // namespace gds
