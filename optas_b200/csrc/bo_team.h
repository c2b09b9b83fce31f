// bo_team.h -- code generation for the team tier (one instance per team of G threads in G warps, state in shared
// memory; csrc/jit/bo_ipm_team.cuh).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "bo_codegen.h"

namespace bo {

// A tape cut into G slices by output.  slices[r] computes the outputs with owner[o] == r (o indexes output_instr,
// the OUTPUT rows of the tape in order); cost[r] is the estimated issue-slot count of slice r.
struct TeamSlices {
  int G = 1;
  std::vector<Tape> slices;
  std::vector<int> owner;
  std::vector<int> out_seg, out_elem;  // segment / element of every output row
  std::vector<int64_t> output_instr;
  std::vector<int64_t> cost;
  int64_t total_cost = 0;
};

TeamSlices slice_tape(const Tape& tape, int G, const std::vector<int>& extra_cost_per_output_segment);

struct TeamPlan {
  int G = 4;
  double rho = 1.0e6;        // weight of the JE'JE term baked into the KX outputs (BO_STATIC_RHO of the kernel)
  uint32_t trig_mask = 0;    // decision variables whose sin / cos are computed once per trial point and shared
  Tape kkt_tape, fc_tape;    // the transformed tapes (see bo_team.cpp) the slices are cut from
  TeamSlices kkt, fc;
};

// false (with the reason) when the problem cannot run on the team tier
bool make_team_plan(const ProblemSource& ps, int G, TeamPlan* plan, std::string* why);
std::string emit_team_source(const ProblemSource& ps, const TeamPlan& plan);
// doubles of shared memory per instance ("elements" of the [element][lane] layout)
int team_smem_elems(const ProblemSource& ps, int G);

}  // namespace bo
