/* ORACLE (test infrastructure): scalar C interpreter of expression tapes -- the CPU stand-in for
 * the CasADi SX virtual machine that evaluates f / df / v / dv inside the reference's solver
 * callbacks (optas/solver.py:716-734 -> casadi Function call).  One instance at a time, plain
 * double arithmetic, libm transcendental functions.  Tape format: include/b200optas.h (bo_tape).
 *
 * Built by oracle/Makefile into oracle/_build/libtape_vm.so; loaded by oracle/tape_vm.py.
 * Never linked into libb200optas.so. */
#include <math.h>
#include <stdint.h>

#include "bo_opcodes.h"

static double apply1(int op, double a) {
  switch (op) {
    case BO_OP_NEG: return -a;
    case BO_OP_SQ: return a * a;
    case BO_OP_SQRT: return sqrt(a);
    case BO_OP_SIN: return sin(a);
    case BO_OP_COS: return cos(a);
    case BO_OP_TAN: return tan(a);
    case BO_OP_ASIN: return asin(a);
    case BO_OP_ACOS: return acos(a);
    case BO_OP_ATAN: return atan(a);
    case BO_OP_FABS: return fabs(a);
    case BO_OP_EXP: return exp(a);
    case BO_OP_LOG: return log(a);
    case BO_OP_NOT: return a == 0.0;
    case BO_OP_SIGN: return (double)((a > 0.0) - (a < 0.0));
    case BO_OP_FLOOR: return floor(a);
    case BO_OP_CEIL: return ceil(a);
    case BO_OP_TANH: return tanh(a);
    case BO_OP_SINH: return sinh(a);
    case BO_OP_COSH: return cosh(a);
  }
  return NAN;
}

static double apply2(int op, double a, double b) {
  switch (op) {
    case BO_OP_ADD: return a + b;
    case BO_OP_SUB: return a - b;
    case BO_OP_MUL: return a * b;
    case BO_OP_DIV: return a / b;
    case BO_OP_ATAN2: return atan2(a, b);
    case BO_OP_FMIN: return fmin(a, b);
    case BO_OP_FMAX: return fmax(a, b);
    case BO_OP_POW: return pow(a, b);
    case BO_OP_LT: return a < b;
    case BO_OP_LE: return a <= b;
    case BO_OP_EQ: return a == b;
    case BO_OP_NE: return a != b;
    case BO_OP_AND: return (a != 0.0) && (b != 0.0);
    case BO_OP_OR: return (a != 0.0) || (b != 0.0);
  }
  return NAN;
}

/* Evaluate one instance.  in[k] / out[k] point at the k-th segment of this instance. */
void tape_eval(const int32_t* instr, int64_t n_instr, const double* consts, double* work,
               const double* const* in, double* const* out) {
  for (int64_t i = 0; i < n_instr; ++i) {
    const int32_t* r = instr + 4 * i;
    const int op = r[0] & 0xFF;
    if (op == BO_OP_INPUT) {
      work[r[1]] = in[r[3]][r[2]];
    } else if (op == BO_OP_CONST) {
      work[r[1]] = consts[r[2]];
    } else if (op == BO_OP_OUTPUT) {
      out[r[3]][r[2]] = work[r[1]];
    } else if (op == BO_OP_IF_ELSE) {
      work[r[1]] = work[(uint32_t)r[0] >> 8] != 0.0 ? work[r[2]] : work[r[3]];
    } else if (op >= BO_OP_NEG) {
      work[r[1]] = apply1(op, work[r[2]]);
    } else {
      work[r[1]] = apply2(op, work[r[2]], work[r[3]]);
    }
  }
}

/* Batched: segment k of instance b lives at in[k] + b * in_sizes[k] (row-major [B][size]). */
void tape_eval_batch(const int32_t* instr, int64_t n_instr, const double* consts, double* work, int64_t B,
                     int32_t n_in, const int32_t* in_sizes, const double* const* in,
                     int32_t n_out, const int32_t* out_sizes, double* const* out) {
  const double* ip[32];
  double* op[32];
  for (int64_t b = 0; b < B; ++b) {
    for (int k = 0; k < n_in; ++k) ip[k] = in[k] + b * in_sizes[k];
    for (int k = 0; k < n_out; ++k) op[k] = out[k] + b * out_sizes[k];
    tape_eval(instr, n_instr, consts, work, ip, op);
  }
}
