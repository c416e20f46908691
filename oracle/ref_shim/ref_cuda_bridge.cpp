// ref_cuda_bridge.cpp — extern "C" doorway into the REFERENCE'S OWN CUDA solver
// (libs/qmcp-solver/src/quasi_mcp_cuda_max_flow_solver.cu, compiled unmodified by nvcc from
// /root/reference into oracle/_ref/libgds_refcuda.so).  TEST INFRASTRUCTURE: lets the GPU tests
// compare our result with quasi-mcp-cuda on the same box, and bench.py time it as a baseline.
#include <cstdint>
#include <cstring>
#include <memory>

#include "bam-api/bam_api.hpp"
#include "qmcp-solver/quasi_mcp_cuda_max_flow_solver.hpp"

extern "C" {

// solve(max_coverage, BamApi(reads)) with the reference's quasi-mcp-cuda.  Writes ascending kept
// read indices, returns their count (or -1 if out_cap is too small).
int64_t ref_cuda_solve(uint64_t n, const uint32_t* s, const uint32_t* e, uint32_t L, uint32_t M,
                       uint64_t* out_ids, uint64_t out_cap) {
    bam_api::AOSPairedReads r;
    r.ref_genome_length = L;
    r.reserve(n);
    for (uint64_t i = 0; i < n; ++i)
        r.push_back(bam_api::Read(i, s[i], e[i], 0, e[i] - s[i] + 1, i % 2 == 0));
    bam_api::BamApi api(r);
    qmcp::QuasiMcpCudaMaxFlowSolver solver;  // a fresh instance per call (state is not reusable, App. B1)
    std::unique_ptr<qmcp::Solution> sol = solver.solve(M, api);
    if (sol->size() > out_cap) return -1;
    std::memcpy(out_ids, sol->data(), sol->size() * sizeof(uint64_t));
    return (int64_t)sol->size();
}
}
