// bo_codegen.h -- expression tape -> straight-line CUDA C++ (host side of libb200optas).
#pragma once
#include <functional>
#include <string>
#include <vector>

#include "b200optas.h"

namespace bo {

// Owning copy of a bo_tape (the C struct only borrows caller memory).
struct Tape {
  std::vector<int32_t> instr;  // [n][4]
  std::vector<double> consts;
  int32_t n_work = 0;
  std::vector<int32_t> in_sizes, out_sizes;
  int64_t n_instr() const { return (int64_t)instr.size() / 4; }
};

struct Sparsity {
  std::vector<int32_t> row, col;
  int32_t nnz() const { return (int32_t)row.size(); }
};

// Validate and copy.  Returns false and fills err on a malformed tape.
bool copy_tape(const bo_tape& in, Tape* out, std::string* err);
bool copy_sparsity(const bo_sparsity& in, int32_t n_rows, int32_t n_cols, bool lower_only, Sparsity* out,
                   std::string* err);

// Emit `BO_DEVICE void <name>(const double* i0, ..., double* o0, ...)` evaluating the tape.
// sin/cos of the same operand are fused into one sincos.
std::string emit_tape_function(const Tape& tape, const std::string& name);
// Same with the function head and the leaf / store statements supplied by the caller (row = the 4-int instruction;
// const_expr may return "" to get the literal).
struct TapeEmitHooks {
  std::string signature, prologue;
  std::function<std::string(int64_t, const int32_t*)> input_expr, const_expr;
  std::function<std::string(int64_t, const int32_t*, const std::string&)> output_stmt;
};
std::string emit_tape_function(const Tape& tape, const std::string& name, const TapeEmitHooks* hooks);

// Count of floating-point operations (adds, muls, ... 1 each; sincos counted as 2 calls).
struct TapeStats {
  int64_t n_arith = 0, n_div_sqrt = 0, n_trig = 0, n_other = 0;
};
TapeStats tape_stats(const Tape& tape);

struct ProblemSource {
  int32_t nx, np, n_eq, n_ineq;
  Tape fc, kkt;
  Sparsity jac_eq, jac_ineq, hess;
};

// Linear equality constraints: y does not enter the Hessian of the Lagrangian, so the least-squares multiplier
// re-estimate after a convexified step (which exists to break the dw -> y -> Hessian -> dw feedback) is skipped.
bool hessian_depends_on_eq_multipliers(const ProblemSource& ps);

// Full translation unit of the tier-S (thread-per-instance) solver for this problem.
struct SparsePlan;
std::string emit_problem_source(const ProblemSource& ps, int threads_per_block, bool pivoted_ldl,
                                const SparsePlan* sparse = nullptr, bool large = false, bool qp = false);
// Quadratic cost, linear constraints: no Jacobian or Hessian output of the kkt tape depends on x, y or z (decided from the
// tape, not from the declared problem class).  Such problems run the dedicated QP iteration (csrc/jit/bo_qp_reg.cuh).
bool problem_is_qp(const ProblemSource& ps);
// Full translation unit of the streaming evaluation kernel for one tape.
std::string emit_function_source(const Tape& tape, int threads_per_block, int out_stages = 2);

}  // namespace bo
