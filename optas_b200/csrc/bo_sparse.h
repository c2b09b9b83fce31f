// bo_sparse.h -- symbolic analysis for the static-order sparse LDL' of the KKT system.
//
// The solver kernel factors K = [ H + JI'S JI + rho JE'JE + dw I , JE' ; JE , -dc I ] without
// pivoting (see csrc/jit/bo_ipm_reg.cuh).  With a fixed elimination order the fill pattern is known
// when the problem is created, so for systems too large to treat densely the host does the
// symbolic work once -- ordering, fill, position maps, an "update program" -- and the kernel runs a
// table-driven numeric factorisation whose control flow is identical for every instance (= every
// lane of a warp).  This replaces what MUMPS redoes inside IPOPT on every solve of the reference
// path (analysis phase).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "bo_codegen.h"

namespace bo {

struct SparsePlan {
  int n = 0;                      // nx + n_eq
  int nx = 0;
  std::vector<int> perm, iperm;   // perm[new] = old, iperm[old] = new
  std::vector<int> colptr, rowidx;  // strict lower triangle of L in the permuted order, column-major
  int nnzL() const { return (int)rowidx.size(); }
  // value array of one instance: vals[0..n) = D (permuted order), vals[n + e] = L entry e
  int vals_size() const { return n + nnzL(); }
  // position of the (old-index) entry (i, j) of the symmetric matrix in vals; -1 if not in the pattern
  int pos(int i_old, int j_old) const;
  // flat table uploaded to the device (layout documented in bo_ipm_reg.cuh, "sparse LDL' tables")
  std::vector<int32_t> table;
  std::vector<double> dtable;     // constants of the interpreted tapes (large mode only)
  int64_t flops = 0;              // multiply-adds of one numeric factorisation
};

// Build the plan from the structural pattern of the KKT matrix implied by the problem.
// Ordering: greedy minimum degree, with a constraint row eligible only once every variable it touches
// has been eliminated (keeps its pivot away from the bare -dc and the order time-interleaved for
// horizon problems).
// `large`: also append the structure tables, the KKT assembly program and the two tapes, so that the
// kernel needs no problem-specific code at all (table-driven everything; see bo_ipm_reg.cuh BO_LARGE).
// `n_segments` > 1: every chain of the KKT graph is cut into that many pieces, separators eliminated last (flat).
// `nd_depth` > 0: nested dissection instead -- the chain is bisected nd_depth times with minimum-vertex-cover separators,
// separators eliminated deepest level first (elimination-tree height ~ piece length + nd_depth separator blocks).
// `nd_bias` in [0.5, 1): where an interval with one free end is cut (0.5 = the middle; larger = nearer its separator end).
SparsePlan make_sparse_plan(const ProblemSource& ps, bool large, int n_segments = 1, int nd_depth = 0, double nd_bias = 0.5);

}  // namespace bo
