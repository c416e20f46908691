// SolverManager — name -> solver registry behind the `-a` switch (src/solver_manager.hpp:16-43).
// Registers the new algorithm next to whatever CPU solvers the build has.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "qmcp-solver/quasi_mcp_b200_max_flow_solver.hpp"
#include "qmcp-solver/solver.hpp"

class SolverManager {
   public:
    SolverManager() {
        solvers_map_.emplace("quasi-mcp-b200", std::make_unique<qmcp::QuasiMcpB200MaxFlowSolver>());
        // the fewest reads that keep min(coverage, M) everywhere: mcp-cpu's objective
        // (mcp_cpu_cost_scaling_solver.cpp:33-67) by the device sweep (gds_params.algorithm = 1)
        solvers_map_.emplace("mcp-b200", std::make_unique<qmcp::McpB200SweepSolver>());
        for (const auto& kv : solvers_map_) algorithms_names_.push_back(kv.first);
    }
    qmcp::Solver& get(const std::string& name) const { return *solvers_map_.at(name); }
    bool contains(const std::string& name) const { return solvers_map_.count(name) != 0; }
    const std::vector<std::string>& get_names() const { return algorithms_names_; }

   private:
    std::map<std::string, std::unique_ptr<qmcp::Solver>> solvers_map_;
    std::vector<std::string> algorithms_names_;
};
