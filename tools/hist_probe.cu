// Probe: how fast can one B200 SM count keys?  Candidates for the direct (sort-free) bundle
// histogram of K2: per read one shared-memory atomic, one global reduction, a mix of both, and
// the (racy, wrong) plain read-modify-write as the upper bound of the shared-memory path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hist_probe hist_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int kKeys = 30016;  // 30 kb sample: one counter per start node (padded)
constexpr int kSMs = 148;

__device__ __forceinline__ uint4 ld4(const uint32_t* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

__global__ void k_gen(uint32_t* S, uint32_t* E, size_t n, uint32_t L, uint32_t R) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t x = i * 0x9E3779B97F4A7C15ull + 0x1234567;
        x ^= x >> 31; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 29; x *= 0x94D049BB133111EBull; x ^= x >> 32;
        uint32_t s = (uint32_t)(x % (L - R + 1));
        S[i] = s;
        E[i] = s + R - 1;
    }
}

enum Mode { ATOMS = 0, RACY, REDG, HYBRID, LOADONLY, ATOMS_RET, HYBRID31 };

// one CTA walks whole samples; counters in shared memory (or the sample's global histogram)
template <int MODE, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS)
k_hist(const uint32_t* __restrict__ S, const uint32_t* __restrict__ E, uint32_t per_sample,
       uint32_t n_samples, uint32_t* __restrict__ ghist, uint32_t R, unsigned long long* sink) {
    extern __shared__ uint32_t h[];
    uint32_t bad = 0, acc = 0;
    for (uint32_t k = blockIdx.x; k < n_samples; k += gridDim.x) {
        for (int i = threadIdx.x; i < kKeys; i += THREADS) h[i] = 0;
        __syncthreads();
        uint32_t* gh = ghist + (size_t)k * kKeys;
        const uint32_t* s = S + (size_t)k * per_sample;
        const uint32_t* e = E + (size_t)k * per_sample;
        const uint32_t n4 = per_sample / 4;
        for (uint32_t j = threadIdx.x; j < n4; j += THREADS * UNROLL) {
            uint4 a[UNROLL], b[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                uint32_t jj = j + u * THREADS;
                if (jj < n4) {
                    a[u] = ld4(s + 4 * (size_t)jj);
                    b[u] = ld4(e + 4 * (size_t)jj);
                } else {
                    a[u] = make_uint4(~0u, ~0u, ~0u, ~0u);
                    b[u] = a[u];
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                uint32_t ks[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
                uint32_t es[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (ks[q] == ~0u) continue;
                    bad += (es[q] - ks[q] + 1 != R) | (ks[q] >= kKeys);
                    if (MODE == ATOMS) atomicAdd(&h[ks[q]], 1u);
                    if (MODE == ATOMS_RET) acc += atomicAdd(&h[ks[q]], 1u);
                    if (MODE == RACY) h[ks[q]] += 1u;
                    if (MODE == REDG) atomicAdd(&gh[ks[q]], 1u);
                    if (MODE == HYBRID) {
                        if (q & 1) atomicAdd(&gh[ks[q]], 1u);
                        else atomicAdd(&h[ks[q]], 1u);
                    }
                    if (MODE == HYBRID31) {
                        if (q == 3) atomicAdd(&gh[ks[q]], 1u);
                        else atomicAdd(&h[ks[q]], 1u);
                    }
                    if (MODE == LOADONLY) acc += ks[q];
                }
            }
        }
        __syncthreads();
        if (MODE != REDG && MODE != LOADONLY) {
            for (int i = threadIdx.x; i < kKeys; i += THREADS) {
                uint32_t c = h[i];
                if (MODE == HYBRID || MODE == HYBRID31) { if (c) atomicAdd(&gh[i], c); }
                else gh[i] = c;
            }
        }
        __syncthreads();
    }
    if (bad | (acc == 0x12345u)) atomicAdd(sink, (unsigned long long)bad + acc);
}

template <int MODE, int THREADS, int UNROLL>
void run(const char* name, const uint32_t* S, const uint32_t* E, uint32_t per_sample, uint32_t ns,
         uint32_t* gh, unsigned long long* sink, int ctas_per_sm) {
    auto kern = k_hist<MODE, THREADS, UNROLL>;
    int smem = kKeys * 4;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemsetAsync(gh, 0, (size_t)ns * kKeys * 4);
        cudaEventRecord(e0);
        kern<<<kSMs * ctas_per_sm, THREADS, smem>>>(S, E, per_sample, ns, gh, 150, sink);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) {
            printf("%s: %s\n", name, cudaGetErrorString(err));
            return;
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    // checksum of the histogram (RACY is expected to lose counts)
    double reads = (double)per_sample * ns;
    double cyc_per_lane = best * 1e-3 * 1.965e9 * kSMs / reads;
    printf("%-28s thr %4d unroll %d ctas/sm %d: %8.3f ms  %7.1f Greads/s  %7.1f GB/s  %.3f SM-cyc/read\n",
           name, THREADS, UNROLL, ctas_per_sm, best, reads / best * 1e-6, reads * 8 / best * 1e-6,
           cyc_per_lane);
    fflush(stdout);
}

__global__ void k_sum(const uint32_t* gh, size_t n, unsigned long long* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long s = 0;
    for (; i < n; i += (size_t)gridDim.x * blockDim.x) s += gh[i];
    atomicAdd(out, s);
}

int main(int argc, char** argv) {
    const uint32_t ns = argc > 1 ? atoi(argv[1]) : 296;  // 2 waves of 148
    const uint32_t per_sample = 2000000;
    const size_t n = (size_t)ns * per_sample;
    uint32_t *S, *E, *gh;
    unsigned long long* sink;
    cudaMalloc(&S, n * 4);
    cudaMalloc(&E, n * 4);
    cudaMalloc(&gh, (size_t)ns * kKeys * 4);
    cudaMalloc(&sink, 16);
    cudaMemset(sink, 0, 16);
    k_gen<<<kSMs * 8, 256>>>(S, E, n, 30000, 150);
    cudaDeviceSynchronize();
    printf("samples %u x %u reads = %.1f M reads, %.2f GB of start+end\n", ns, per_sample, n * 1e-6,
           n * 8e-9);
    run<LOADONLY, 1024, 2>("load only", S, E, per_sample, ns, gh, sink, 1);
    run<LOADONLY, 1024, 4>("load only", S, E, per_sample, ns, gh, sink, 1);
    run<LOADONLY, 512, 4>("load only", S, E, per_sample, ns, gh, sink, 1);
    run<ATOMS, 1024, 2>("ATOMS", S, E, per_sample, ns, gh, sink, 1);
    run<ATOMS, 1024, 1>("ATOMS", S, E, per_sample, ns, gh, sink, 1);
    run<ATOMS, 512, 2>("ATOMS", S, E, per_sample, ns, gh, sink, 1);
    run<ATOMS, 256, 4>("ATOMS", S, E, per_sample, ns, gh, sink, 1);
    {
        unsigned long long hs = 0;
        cudaMemset(sink + 1, 0, 8);
        k_sum<<<kSMs * 4, 256>>>(gh, (size_t)ns * kKeys, sink + 1);
        cudaMemcpy(&hs, sink + 1, 8, cudaMemcpyDeviceToHost);
        printf("  ATOMS histogram total %llu (expect %zu)\n", hs, n);
    }
    run<ATOMS_RET, 1024, 2>("ATOMS with return", S, E, per_sample, ns, gh, sink, 1);
    run<RACY, 1024, 2>("racy LDS/STS (upper bound)", S, E, per_sample, ns, gh, sink, 1);
    run<REDG, 1024, 2>("REDG to L2", S, E, per_sample, ns, gh, sink, 1);
    run<REDG, 1024, 4>("REDG to L2", S, E, per_sample, ns, gh, sink, 1);
    run<REDG, 512, 4>("REDG to L2", S, E, per_sample, ns, gh, sink, 2);
    run<HYBRID, 1024, 2>("hybrid ATOMS+REDG 1:1", S, E, per_sample, ns, gh, sink, 1);
    {
        unsigned long long hs = 0;
        cudaMemset(sink + 1, 0, 8);
        k_sum<<<kSMs * 4, 256>>>(gh, (size_t)ns * kKeys, sink + 1);
        cudaMemcpy(&hs, sink + 1, 8, cudaMemcpyDeviceToHost);
        printf("  hybrid histogram total %llu (expect %zu)\n", hs, n);
    }
    run<HYBRID31, 1024, 2>("hybrid ATOMS+REDG 3:1", S, E, per_sample, ns, gh, sink, 1);
    unsigned long long hsink = 0;
    cudaMemcpy(&hsink, sink, 8, cudaMemcpyDeviceToHost);
    printf("sink %llu (0 = every read passed the checks)\n", hsink);
    return 0;
}
