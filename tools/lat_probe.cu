// lat_probe.cu — cost of the primitives a K3 step is made of, in SM clocks, on the box at hand.
// One CTA; thread 0 reports.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void q_append(uint32_t* q, uint32_t* count, uint32_t v) {
    auto g = cg::coalesced_threads();
    uint32_t base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(count, g.size());
    base = g.shfl(base, 0);
    q[(base + g.thread_rank()) & 1023] = v;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_probe(uint32_t* chain, uint32_t n_chain, long long* out) {
    extern __shared__ uint32_t dyn[];
    if (threadIdx.x == 0 && n_chain == 0xffffffffu) dyn[0] = 1;
    __shared__ uint32_t sm[4096];
    __shared__ uint32_t cnt, flag[3];
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < 4096; i += THREADS) sm[i] = (i * 7 + 1) & 4095;
    if (tid == 0) { cnt = 0; flag[0] = flag[1] = flag[2] = 0; }
    __syncthreads();
    const int R = 200;
    long long t0, t1;
    uint32_t x = tid;
    // (0) __syncthreads
    t0 = clock64();
    for (int r = 0; r < R; ++r) __syncthreads();
    t1 = clock64();
    if (tid == 0) out[0] = (t1 - t0) / R;
    // (1) __syncthreads_or
    int acc = 0;
    t0 = clock64();
    for (int r = 0; r < R; ++r) acc += __syncthreads_or(x == (uint32_t)r);
    t1 = clock64();
    if (tid == 0) out[1] = (t1 - t0) / R + (acc == 12345);
    // (2) flag + __syncthreads (rotating flags)
    t0 = clock64();
    for (int r = 0; r < R; ++r) {
        if (x == (uint32_t)r) flag[r % 3] = 1;
        if (tid == 0) flag[(r + 1) % 3] = 0;
        __syncthreads();
        acc += flag[r % 3];
    }
    t1 = clock64();
    if (tid == 0) out[2] = (t1 - t0) / R + (acc == 12345);
    // (3) dependent LDS chain (thread 0 only)
    if (tid == 0) {
        uint32_t p = 0;
        t0 = clock64();
        for (int r = 0; r < R; ++r) p = sm[p];
        t1 = clock64();
        out[3] = (t1 - t0) / R + (p == 99999);
    }
    __syncthreads();
    // (4) dependent smem atomicAdd with return (thread 0)
    if (tid == 0) {
        uint32_t p = 0;
        t0 = clock64();
        for (int r = 0; r < R; ++r) p = atomicAdd(&sm[p & 4095], 1u) & 4095;
        t1 = clock64();
        out[4] = (t1 - t0) / R + (p == 99999);
    }
    __syncthreads();
    // (5) smem atomicCAS dependent (thread 0)
    if (tid == 0) {
        uint32_t p = 0;
        t0 = clock64();
        for (int r = 0; r < R; ++r) p = atomicCAS(&sm[p & 4095], 77u, 78u) & 4095;
        t1 = clock64();
        out[5] = (t1 - t0) / R + (p == 99999);
    }
    __syncthreads();
    // (6) coalesced append by 5 lanes of warp 0, dependent on previous value
    if (tid < 5) {
        uint32_t p = tid;
        t0 = clock64();
        for (int r = 0; r < R; ++r) { q_append(sm, &cnt, p); p = sm[(p + r) & 1023]; }
        t1 = clock64();
        if (tid == 0) out[6] = (t1 - t0) / R + (p == 99999);
    }
    __syncthreads();
    // (7) dependent global loads, random chain over n_chain 32-byte records (L2 resident after warm-up)
    if (tid == 0) {
        uint32_t p = 0;
        for (int r = 0; r < 4096; ++r) p = __ldcg(&chain[p * 8]);  // warm: the cycle has 4096 records
        p = 0;
        t0 = clock64();
        for (int r = 0; r < R; ++r) p = chain[p * 8];
        t1 = clock64();
        out[7] = (t1 - t0) / R + (p == 0xffffffffu);
        // (8) the same with ld.global.cg (L2 only)
        t0 = clock64();
        for (int r = 0; r < R; ++r) p = __ldcg(&chain[p * 8]);
        t1 = clock64();
        out[8] = (t1 - t0) / R + (p == 0xffffffffu);
        // (9) store then dependent load of the same address (write-through + L1 behaviour)
        t0 = clock64();
        for (int r = 0; r < R; ++r) { chain[p * 8 + 1] = r; p = chain[p * 8]; }
        t1 = clock64();
        out[9] = (t1 - t0) / R + (p == 0xffffffffu);
        // (10) global atomicAdd with return, dependent
        t0 = clock64();
        for (int r = 0; r < R; ++r) p = atomicAdd(&chain[p * 8 + 2], 0u) * 0 + chain[p * 8];
        t1 = clock64();
        out[10] = (t1 - t0) / R + (p == 0xffffffffu);
    }
    __syncthreads();
    // (11) ATOMS.OR fire-and-forget x2 + LDS per iteration (all threads), then barrier
    t0 = clock64();
    for (int r = 0; r < R; ++r) {
        const uint32_t a = sm[(x + r) & 4095];
        atomicOr(&sm[(a + 1) & 4095], 1u << (r & 31));
        atomicOr(&sm[(a + 150) & 4095], 1u << (r & 31));
        __syncthreads();
    }
    t1 = clock64();
    if (tid == 0) out[11] = (t1 - t0) / R;
    // (13)/(14) one dependent L2 load per iteration then a CTA barrier, without / with a global store
    // before the barrier: does BAR.SYNC wait for the store?
    {
        uint32_t p = tid == 0 ? 0u : tid;
        t0 = clock64();
        for (int r = 0; r < R; ++r) {
            if (tid < 4) p = __ldcg(&chain[(p & (n_chain - 1)) * 8]);
            __syncthreads();
        }
        t1 = clock64();
        if (tid == 0) out[13] = (t1 - t0) / R + (p == 0xffffffffu);
        t0 = clock64();
        for (int r = 0; r < R; ++r) {
            if (tid < 4) {
                p = __ldcg(&chain[(p & (n_chain - 1)) * 8]);
                chain[(p & (n_chain - 1)) * 8 + 3] = r;
            }
            __syncthreads();
        }
        t1 = clock64();
        if (tid == 0) out[14] = (t1 - t0) / R + (p == 0xffffffffu);
        // (15) store, barrier, then ANOTHER thread loads what was stored (plain load through L1)
        t0 = clock64();
        for (int r = 0; r < R; ++r) {
            if (tid == 0) chain[(r & 1023) * 8 + 4] = r + 1;
            __syncthreads();
            if (tid == 32) p += chain[(r & 1023) * 8 + 4];
            __syncthreads();
        }
        t1 = clock64();
        if (tid == 32) out[15] = (t1 - t0) / R + (p == 0xffffffffu);
    }
    __syncthreads();
    // (16) touch a record (load nobody waits for), do ~1500 clocks of other work, then load it: L1 hit?
    if (tid == 0) {
        uint32_t p = 5, acc2 = 0;
        long long tot = 0;
        for (int r = 0; r < R; ++r) {
            const uint32_t nxt = (p * 2654435761u >> 8) & (n_chain - 1);
            uint32_t dummy;
            asm volatile("ld.global.u32 %0, [%1];" : "=r"(dummy) : "l"(chain + (size_t)nxt * 8));
            for (int k = 0; k < 60; ++k) acc2 = sm[(acc2 + k) & 4095];  // ~1700 clocks of LDS chain
            const long long a = clock64();
            uint32_t v;
            asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(chain + (size_t)nxt * 8 + (acc2 & 0)) : "memory");
            p = v + acc2 * 0 + nxt;
            tot += clock64() - a + (p == 0xffffffffu);
        }
        out[16] = tot / R;
        // (17) the same without the touch
        tot = 0;
        for (int r = 0; r < R; ++r) {
            const uint32_t nxt = (p * 2654435761u >> 8) & (n_chain - 1);
            for (int k = 0; k < 60; ++k) acc2 = sm[(acc2 + k) & 4095];
            const long long a = clock64();
            uint32_t v;
            asm volatile("ld.global.u32 %0, [%1];" : "=r"(v) : "l"(chain + (size_t)nxt * 8 + (acc2 & 0)) : "memory");
            p = v + acc2 * 0 + nxt;
            tot += clock64() - a + (p == 0xffffffffu);
        }
        out[17] = tot / R;
    }
    __syncthreads();
    // (12) first touch of a cold line chain (DRAM): separate region never touched
    if (tid == 0) {
        uint32_t p = n_chain;  // second half of the buffer, cold
        t0 = clock64();
        for (int r = 0; r < 50; ++r) p = chain[(size_t)p * 8] + n_chain;
        t1 = clock64();
        out[12] = (t1 - t0) / 50 + (p == 0xffffffffu);
    }
}

int main() {
    const uint32_t n = 1 << 12;  // 4 Ki records x 32 B = 128 KB: L2 resident, larger than the L1 left over
    uint32_t* h = (uint32_t*)malloc((size_t)2 * n * 32);
    // random cyclic permutation
    uint32_t* perm = (uint32_t*)malloc(n * 4);
    for (uint32_t i = 0; i < n; ++i) perm[i] = i;
    uint64_t s = 88172645463325252ull;
    for (uint32_t i = n - 1; i > 0; --i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        uint32_t j = s % (i + 1);
        uint32_t t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
    for (int half = 0; half < 2; ++half)
        for (uint32_t i = 0; i < n; ++i) {
            uint32_t* rec = h + ((size_t)half * n + perm[i]) * 8;
            for (int k = 0; k < 8; ++k) rec[k] = 0;
            rec[0] = perm[(i + 1) % n];
        }
    uint32_t* d;
    long long* out;
    cudaMalloc(&d, (size_t)2 * n * 32);
    cudaMalloc(&out, 32 * 8);
    cudaMemcpy(d, h, (size_t)2 * n * 32, cudaMemcpyHostToDevice);
    // evict L2: touch a big buffer
    char* big;
    cudaMalloc(&big, 512u << 20);
    cudaMemset(big, 1, 512u << 20);
    const char* names[18] = {"__syncthreads", "__syncthreads_or", "flag+__syncthreads", "LDS dependent",
                             "ATOMS.ADD ret dependent", "ATOMS.CAS dependent", "coalesced append (5 lanes)+LDS",
                             "LDG dependent (L2 hit, via L1)", "LDG.cg dependent (L2)", "STG + LDG dependent",
                             "ATOMG ret + LDG", "LDS + 2 ATOMS.OR + barrier", "LDG cold (DRAM)", "LDG.cg + barrier", "LDG.cg + STG + barrier", "STG, barrier, LDG by another warp, barrier", "LDG 1700 clocks after a touch", "LDG without the touch"};
    long long ho[32];
    for (int pass = 0; pass < 2; ++pass) {
        cudaMemset(out, 0, 32 * 8);
        if (pass == 0) k_probe<256><<<1, 256>>>(d, n, out);
        else {
            cudaFuncSetAttribute(k_probe<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            k_probe<256><<<1, 256, 200 * 1024>>>(d, n, out);
        }
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(ho, out, 32 * 8, cudaMemcpyDeviceToHost);
        printf("---- 256 threads, %s ----\n", pass == 0 ? "no dynamic shared memory" : "200 KB dynamic shared memory");
        for (int i = 0; i < 18; ++i) printf("%-36s %6lld clocks\n", names[i], ho[i]);
        cudaMemset(big, 2, 512u << 20);
    }
    return 0;
}
