// bo_coop.h -- host-side plan of the COOPERATIVE tier: one problem instance per CTA.
//
// For horizon problems (C3/C4/C5: KKT systems of 10^2..10^3 rows, tapes of 10^3..10^5 instructions) one
// thread per instance is the wrong shape: the per-instance state (factor, vectors, tape work array) is
// 10^4..10^5 bytes, far beyond registers and L1, and the iteration is a chain of dependent L2 accesses.
// Here a whole CTA works on one instance:
//   * the sparse LDL' factor lives in SHARED memory (C4: 183 KB of the 227 KB) and is computed by a
//     level-scheduled, target-owned (deterministic, atomics-free) left-looking program;
//   * the expression tapes are PARTITIONED into independent sub-tapes (one per horizon stage for the
//     workloads here), one per thread; outputs that sum over several sub-tapes (the cost) are reduced
//     from per-thread partials in a fixed order;
//   * sub-expressions that depend only on the parameters are evaluated once per instance, not once
//     per iteration;
//   * vector work (Jacobian products, KKT assembly, line-search measures) runs as thread-parallel
//     gathers + block reductions.
// Everything is table-driven: this file builds the tables once per problem (bo_problem_create); the
// kernel is csrc/jit/bo_ipm_cta.cuh.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "bo_sparse.h"

namespace bo {

// header slots of CoopPlan::itab (all offsets are absolute int32 indices into itab)
enum CoopSlot {
  CT_N = 0, CT_NNZL, CT_PERM, CT_SIGN, CT_NLEV, CT_PROG_FAC, CT_PROG_FWD, CT_PROG_BWD,
  CT_ANT, CT_APOS, CT_APTR, CT_ATERM,
  CT_JE_RPTR, CT_JE_RENT, CT_JE_CPTR, CT_JE_CENT,
  CT_JI_RPTR, CT_JI_RENT, CT_JI_CPTR, CT_JI_CENT,
  CT_TAPE_FC, CT_TAPE_KKT, CT_HEADER = 64
};
// slots of a tape section (relative to the section start; offsets inside are absolute)
enum CoopTapeSlot {
  TS_NSUB = 0, TS_CONST0, TS_NPE, TS_NPART, TS_PRE_N, TS_PRE_OFF, TS_LEN_OFF, TS_WBASE_OFF, TS_STREAM_OFF,
  TS_NRED, TS_RED_OFF, TS_PART_SEG, TS_PE_SEG, TS_GEN_NWARP, TS_GEN_WL_OFF, TS_HEADER = 16
};

struct CoopTapeInfo {
  int nsub = 0, n_pe = 0, n_part = 0, n_red = 0, n_components = 0, n_split_outputs = 0;
  int64_t total_instr = 0, max_len = 0, pre_len = 0;
  int n_classes = 0;              // generated tapes: component classes (one straight-line function each)
  int64_t code_rows = 0;          // generated tapes: instructions over all class functions
  int64_t warp_rows = 0;          // generated tapes: longest per-warp work list (instructions executed in lock step)
};

struct CoopPlan {
  SparsePlan sp;               // ordering + fill pattern (its own device tables are not used by this tier)
  std::vector<int32_t> itab;   // all integer tables
  std::vector<double> dtab;    // constants of the two tapes
  int n_work_fc = 1, n_work_kkt = 1, n_work_pre = 1;  // interpreter work slots per sub-tape
  int fc_wstride = 0, kkt_wstride = 0;  // stride of the shared-memory work arrays w[slot][thread]; 0 = thread-local
  int smem_doubles = 0;        // dynamic shared memory of the kernel
  bool w_in_smem = false;      // the per-instance vector workspace (coop_scratch_doubles) lives in shared memory too, after
                               // those smem_doubles (small problems: C3); else in a per-CTA slice of global memory
  bool gen_tapes = false;      // tapes compiled to straight-line code per component class (else interpreted)
  std::string gen_code;        // the generated class functions + dispatchers
  int n_levels = 0;            // elimination-tree height (= barriers per factorisation)
  int n_segments = 1;          // elimination order: 1 = minimum degree, k > 1 = chains cut flat into k pieces, -d = nested dissection of depth d
  int ldl_g = 1, ldl_w = 1;    // factor program: lanes per target, participating warps
  bool fac_aligned = false;    // factor program in the column-aligned form (bo_coop.cpp)
  int fac_pk = 8, fwd_pk = 2, bwd_pk = 2;  // operand words per packet of the factor / substitution lane programs
  int solve_g = 1, solve_bwd_g = 1;  // lanes cooperating on one row (forward) / column (backward) of a triangular solve
  int64_t fac_steps = 0, solve_steps = 0;  // longest warp stream of the factor / both triangular solves
  double asm_terms_per_position = 0.0;  // KKT assembly: mean number of terms per assembled position
  double prog_cost = 0.0;      // predicted cycles of one factorisation + one pair of substitutions (scheduler's cost model)
  int64_t n_contrib = 0;       // multiply-adds of one numeric factorisation
  CoopTapeInfo fc, kkt;
  int vals_size() const { return sp.vals_size(); }
};

// `cache_dir` (optional): where the verdict of the elimination-order search is remembered across processes.
CoopPlan make_coop_plan(const ProblemSource& ps, int threads_per_block, const std::string& cache_dir = std::string());

// Doubles of per-CTA global workspace (must equal BO_SCRATCH_DOUBLES of csrc/jit/bo_ipm_cta.cuh).
size_t coop_scratch_doubles(const ProblemSource& ps, const CoopPlan& plan);

// Prelude (sizes only -- the tier has no problem-specific code) + #include of the kernel.
std::string emit_coop_source(const ProblemSource& ps, const CoopPlan& plan, int threads_per_block);

}  // namespace bo
