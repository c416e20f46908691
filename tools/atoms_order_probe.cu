// Probe: when several lanes of ONE warp instruction do atomicAdd (with return) on the SAME
// shared-memory word, in which order does the hardware serialise them?  If the returned values
// always ascend with the lane id, a shared-memory atomic is a STABLE rank within the instruction.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms_order_probe atoms_order_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(uint32_t seed, int iters, int nbins, unsigned long long* viol, unsigned long long* conflicts) {
    __shared__ uint32_t cnt[32][256];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t x = seed ^ (blockIdx.x * 9781u + threadIdx.x * 6271u + 1u);
    unsigned long long v = 0, c = 0;
    for (int it = 0; it < iters; ++it) {
        for (int i = lane; i < 256; i += 32) cnt[warp][i] = 0;
        __syncwarp();
        x = x * 1664525u + 1013904223u;
        uint32_t d = (x >> 13) % nbins;
        // half of the iterations run with some lanes inactive (divergent callers)
        bool active = (it & 1) ? ((x >> 5) & 3) != 0 : true;
        uint32_t r = 0xffffffffu;
        if (active) r = atomicAdd(&cnt[warp][d], 1u);
        __syncwarp();
        uint32_t am = __ballot_sync(0xffffffffu, active);
        uint32_t peers = __match_any_sync(0xffffffffu, active ? d : 1000u + lane) & am;
        if (active) {
            uint32_t expect = __popc(peers & ((1u << lane) - 1));
            if (__popc(peers) > 1) ++c;
            if (r != expect) ++v;
        }
        __syncwarp();
    }
    atomicAdd(viol, v);
    atomicAdd(conflicts, c);
}

int main() {
    unsigned long long *dv, hv[2];
    cudaMalloc(&dv, 16);
    for (int nbins : {256, 128, 16, 3, 1}) {
        cudaMemset(dv, 0, 16);
        probe<<<148 * 2, 1024>>>(12345u + nbins, 2000, nbins, dv, dv + 1);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hv, dv, 16, cudaMemcpyDeviceToHost);
        printf("nbins %3d: lanes in conflict %llu, order violations %llu (%s)\n", nbins, hv[1], hv[0], cudaGetErrorString(e));
    }
    return 0;
}
