// qmcp::Solver — the plugin interface every algorithm implements, identical in names, argument
// meaning and ownership to libs/qmcp-solver/include/qmcp-solver/solver.hpp:13-20.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "bam-api/bam_api.hpp"

namespace qmcp {

// ascending indices into BamApi's (post-filter) read arrays
typedef std::vector<bam_api::ReadIndex> Solution;

class Solver {
   public:
    virtual ~Solver() = default;
    // borrowed BamApi, caller owns the returned Solution
    virtual std::unique_ptr<Solution> solve(uint32_t max_coverage, bam_api::BamApi& bam_api) = 0;
    // false => App selects AmpliconBehaviour::FILTER (src/app.cpp:120-128)
    virtual bool uses_quality_of_reads() = 0;
};

}  // namespace qmcp
