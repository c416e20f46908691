// bfs_probe.cu — the first relabel of K3 in isolation: reverse BFS over a 30 001-node line with one
// in-arc per node (u-150 -> u) and the back arcs (u+1 -> u), labels in shared memory.  Variants of
// the level loop, clocks per level.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr uint32_t kInf16 = 0xffffu;
constexpr int N = 30001, R = 150;
constexpr int W = (N + 31) / 32, W4 = (W + 3) & ~3;

__device__ __forceinline__ void mark(uint32_t* word, uint32_t* nxt, uint32_t x) {
    if ((word[x] & 0xffffu) == kInf16) atomicOr(&nxt[x >> 5], 1u << (x & 31));
}

// variant 0: bitmap, 128-bit groups, one group per thread per iteration
// variant 1: bitmap, but the thread -> group map is interleaved by warp (group j handled by lane-major order)
// variant 2: queue with CAS claim and plain atomicAdd append
template <int THREADS, int VARIANT>
__global__ void __launch_bounds__(THREADS, 1) k_bfs(long long* out, uint32_t* levels_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* word = reinterpret_cast<uint32_t*>(smem);
    uint32_t* bmA = word + ((N + 3) & ~3);
    uint32_t* bmB = bmA + W4;
    uint32_t* qa = bmB + W4;
    uint32_t* qb = qa + 4096;
    __shared__ uint32_t flag[3], cnt[3];
    const uint32_t tid = threadIdx.x;
    for (int rep = 0; rep < 3; ++rep) {
        for (uint32_t v = tid; v < N; v += THREADS)
            word[v] = ((v >= R ? v - R : kInf16) << 16) | kInf16;
        for (uint32_t i = tid; i < 2 * W4; i += THREADS) bmA[i] = 0;
        if (tid < 3) { flag[tid] = 0; cnt[tid] = 0; }
        __syncthreads();
        if (VARIANT == 4 && tid == 0) { bmA[(N - 1) >> 5] |= 1u << ((N - 1) & 31); bmA[(N - 2) >> 5] |= 1u << ((N - 2) & 31); }
        if (VARIANT < 2) {
            if (tid == 0) { bmA[(N - 1) >> 5] |= 1u << ((N - 1) & 31); bmA[(N - 2) >> 5] |= 1u << ((N - 2) & 31); }
        } else {
            if (tid == 0) { word[N - 1] = (word[N - 1] & 0xffff0000u) | 1; word[N - 2] = (word[N - 2] & 0xffff0000u) | 1; qa[0] = N - 1; qa[1] = N - 2; cnt[0] = 2; }
        }
        __syncthreads();
        const long long t0 = clock64();
        uint32_t level = 1;
        if (VARIANT < 2) {
            uint32_t* cur = bmA;
            uint32_t* nxt = bmB;
            for (;;) {
                bool any = false;
                if (tid == 0) flag[(level + 1) % 3] = 0;
#pragma unroll 1
                for (uint32_t j = tid; j < W4 / 4; j += THREADS) {
                    uint4* grp = reinterpret_cast<uint4*>(cur) + j;
                    const uint4 q = *grp;
                    if ((q.x | q.y | q.z | q.w) == 0) continue;
                    *grp = make_uint4(0, 0, 0, 0);
#pragma unroll 1
                    for (uint32_t c4 = 0; c4 < 4; ++c4) {
                        uint32_t bits = c4 == 0 ? q.x : c4 == 1 ? q.y : c4 == 2 ? q.z : q.w;
#pragma unroll 1
                        while (bits) {
                            const uint32_t u = (j * 4 + c4) * 32 + (__ffs(bits) - 1);
                            bits &= bits - 1;
                            const uint32_t wu = word[u];
                            if ((wu & 0xffffu) != kInf16) continue;
                            any = true;
                            reinterpret_cast<uint16_t*>(word)[2 * u] = (uint16_t)level;
                            const uint32_t code = wu >> 16;
                            if (VARIANT == 1) {
                                if (u + 1 < N) atomicOr(&nxt[(u + 1) >> 5], 1u << ((u + 1) & 31));
                                if (code < 0xfffeu) atomicOr(&nxt[code >> 5], 1u << (code & 31));
                            } else {
                                if (u + 1 < N) mark(word, nxt, u + 1);
                                if (code < 0xfffeu) mark(word, nxt, code);
                            }
                        }
                    }
                }
                if (any) flag[level % 3] = 1;
                __syncthreads();
                if (!flag[level % 3]) break;
                uint32_t* t = cur; cur = nxt; nxt = t;
                ++level;
            }
        } else {
            uint32_t* T = qa;
            uint32_t* Nq = qb;
            for (;;) {
                const uint32_t c = cnt[(level - 1) % 3];
                if (c == 0) break;
                const uint32_t nl = level + 1;
                uint32_t* nx = &cnt[level % 3];
                if (tid == 0) cnt[(level + 1) % 3] = 0;
#pragma unroll 1
                for (uint32_t i = tid; i < c; i += THREADS) {
                    const uint32_t w = T[i];
                    const uint32_t code = word[w] >> 16;
                    auto claim = [&](uint32_t u) {
                        uint32_t old = word[u];
                        for (;;) {
                            if ((old & 0xffffu) != kInf16) return false;
                            const uint32_t seen = atomicCAS(word + u, old, (old & 0xffff0000u) | nl);
                            if (seen == old) return true;
                            old = seen;
                        }
                    };
                    auto claim_or = [&](uint32_t u) {  // visited bitmap in bmA: one atomic with return
                        const uint32_t bit = 1u << (u & 31);
                        if (atomicOr(&bmA[u >> 5], bit) & bit) return false;
                        reinterpret_cast<uint16_t*>(word)[2 * u] = (uint16_t)nl;
                        return true;
                    };
                    auto append = [&](uint32_t v) {
                        if (VARIANT == 2) {
                            Nq[atomicAdd(nx, 1u)] = v;
                        } else {  // aggregated over the lanes that are here together
                            const uint32_t m = __activemask();
                            const uint32_t lane = threadIdx.x & 31;
                            const int leader = __ffs(m) - 1;
                            uint32_t base = 0;
                            if ((int)lane == leader) base = atomicAdd(nx, __popc(m));
                            base = __shfl_sync(m, base, leader);
                            Nq[base + __popc(m & ((1u << lane) - 1))] = v;
                        }
                    };
                    if (VARIANT == 4) {
                        const bool a = w + 1 < N && claim_or(w + 1);
                        const bool b = code < 0xfffeu && claim_or(code);
                        if (a) append(w + 1);
                        if (b) append(code);
                    } else {
                        if (w + 1 < N && claim(w + 1)) append(w + 1);
                        if (code < 0xfffeu && claim(code)) append(code);
                    }
                }
                __syncthreads();
                uint32_t* t = T; T = Nq; Nq = t;
                ++level;
            }
        }
        const long long t1 = clock64();
        if (tid == 0) { out[rep] = t1 - t0; levels_out[rep] = level; }
        __syncthreads();
    }
}

template <int THREADS, int VARIANT>
void run(const char* name, long long* out, uint32_t* lv) {
    const int smem = 4 * ((N + 3) & ~3) + 8 * W4 + 2 * 4096 * 4;
    cudaFuncSetAttribute(k_bfs<THREADS, VARIANT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_bfs<THREADS, VARIANT><<<1, THREADS, smem>>>(out, lv);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[3]; uint32_t hl[3];
    cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
    cudaMemcpy(hl, lv, 12, cudaMemcpyDeviceToHost);
    printf("%-34s %4d threads: %s levels %u clocks %lld %lld %lld -> %lld per level\n", name, THREADS,
           e == cudaSuccess ? "ok" : cudaGetErrorString(e), hl[2], h[0], h[1], h[2], h[2] / (hl[2] ? hl[2] : 1));
}

int main() {
    long long* out; uint32_t* lv;
    cudaMalloc(&out, 64); cudaMalloc(&lv, 64);
    run<256, 0>("bitmap 128-bit groups", out, lv);
    run<512, 0>("bitmap 128-bit groups", out, lv);
    run<1024, 0>("bitmap 128-bit groups", out, lv);
    run<256, 1>("bitmap, unconditional marks", out, lv);
    run<256, 2>("queue CAS + atomicAdd append", out, lv);
    run<512, 2>("queue CAS + atomicAdd append", out, lv);
    run<256, 3>("queue CAS + aggregated append", out, lv);
    run<256, 4>("queue OR-claim + aggregated append", out, lv);
    run<512, 4>("queue OR-claim + aggregated append", out, lv);
    run<128, 4>("queue OR-claim + aggregated append", out, lv);
    return 0;
}
