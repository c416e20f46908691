// common.cuh — shared helpers for the sm_100a kernels of libgds_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

namespace gds {

constexpr uint32_t kLabelInf = 0x3fffffffu;
constexpr int kNumSMs = 148;  // B200

struct CudaFail {
    cudaError_t err;
    const char* what;
    const char* file;
    int line;
};

#define GDS_CUDA(call)                                                   \
    do {                                                                 \
        cudaError_t _e = (call);                                         \
        if (_e != cudaSuccess) throw gds::CudaFail{_e, #call, __FILE__, __LINE__}; \
    } while (0)

#define GDS_KERNEL_CHECK() GDS_CUDA(cudaGetLastError())

// Grow-only device buffer (arena slot).  Reused across gds_solve calls.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T>
    T* get(size_t n) {
        size_t bytes = n * sizeof(T);
        if (bytes == 0) bytes = sizeof(T);
        if (bytes > cap) {
            if (p) cudaFree(p);
            p = nullptr;
            size_t want = bytes + bytes / 8 + 256;
            cudaError_t e = cudaMalloc(&p, want);
            if (e != cudaSuccess) {
                cap = 0;
                p = nullptr;
                throw CudaFail{e, "cudaMalloc", __FILE__, __LINE__};
            }
            cap = want;
        }
        return reinterpret_cast<T*>(p);
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// Warp-aggregated append: every calling lane with pred==true gets a distinct slot of *counter.
// Must be called by all 32 lanes of the warp (converged).
__device__ __forceinline__ uint32_t warp_agg_slot(bool pred, uint32_t* counter) {
    uint32_t mask = __ballot_sync(0xffffffffu, pred);
    uint32_t base = 0;
    if (mask != 0) {
        int leader = __ffs(mask) - 1;
        if ((int)lane_id() == leader) base = atomicAdd(counter, __popc(mask));
        base = __shfl_sync(0xffffffffu, base, leader);
    }
    return base + __popc(mask & lanemask_lt());
}

// streaming loads that do not pollute L1 (inputs are read once)
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_stream4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

inline int div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
inline int bits_for(uint64_t max_value) {  // bits needed to represent values 0..max_value
    int b = 0;
    while (max_value) {
        ++b;
        max_value >>= 1;
    }
    return b;
}

}  // namespace gds
