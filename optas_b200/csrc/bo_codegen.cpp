// bo_codegen.cpp -- expression tape -> straight-line CUDA C++.
//
// The reference evaluates its cost / constraint / kinematics graphs by interpreting CasADi's
// instruction tape on the CPU inside nlpsol (optas/solver.py:395).  Here the same kind of tape is
// turned into straight-line device code once per problem (at bo_problem_create), so that on the
// GPU every instance's evaluation lives in registers and the compiler can schedule, fuse FMAs and
// fold constants (structural +-1/0 entries of joint-limit Jacobians disappear entirely).
#include "bo_codegen.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <map>
#include <sstream>

#include "bo_opcodes.h"
#include "bo_sparse.h"

namespace bo {

static bool is_unary(int op) { return op >= BO_OP_NEG && op <= BO_OP_COSH; }
static bool is_binary(int op) { return op >= BO_OP_ADD && op <= BO_OP_OR; }

bool copy_tape(const bo_tape& in, Tape* out, std::string* err) {
  auto fail = [&](const std::string& m) {
    *err = "tape: " + m;
    return false;
  };
  if (in.n_instr < 0 || in.n_consts < 0 || in.n_work <= 0 || in.n_in < 0 || in.n_out < 0) return fail("negative size");
  if ((in.n_instr && !in.instr) || (in.n_consts && !in.consts) || (in.n_in && !in.in_sizes) || (in.n_out && !in.out_sizes))
    return fail("null array");
  out->instr.assign(in.instr, in.instr + 4 * in.n_instr);
  out->consts.assign(in.consts, in.consts + in.n_consts);
  out->n_work = in.n_work;
  out->in_sizes.assign(in.in_sizes, in.in_sizes + in.n_in);
  out->out_sizes.assign(in.out_sizes, in.out_sizes + in.n_out);
  for (int32_t s : out->in_sizes)
    if (s < 0) return fail("negative input size");
  for (int32_t s : out->out_sizes)
    if (s < 0) return fail("negative output size");
  std::vector<char> defined(in.n_work, 0);
  for (int64_t i = 0; i < in.n_instr; ++i) {
    const int32_t* r = &out->instr[4 * i];
    const int op = r[0] & 0xFF, c = (int)((uint32_t)r[0] >> 8);
    const int dst = r[1], a = r[2], b = r[3];
    if (dst < 0 || dst >= in.n_work) return fail("work index out of range");
    auto need = [&](int s) { return s >= 0 && s < in.n_work && defined[s]; };
    if (op == BO_OP_INPUT) {
      if (b < 0 || b >= in.n_in || a < 0 || a >= out->in_sizes[b]) return fail("input reference out of range");
      defined[dst] = 1;
    } else if (op == BO_OP_CONST) {
      if (a < 0 || a >= in.n_consts) return fail("constant index out of range");
      defined[dst] = 1;
    } else if (op == BO_OP_OUTPUT) {
      if (b < 0 || b >= in.n_out || a < 0 || a >= out->out_sizes[b]) return fail("output reference out of range");
      if (!need(dst)) return fail("output of an undefined value");
    } else if (op == BO_OP_IF_ELSE) {
      if (!need(a) || !need(b) || !need(c)) return fail("if_else operand undefined");
      defined[dst] = 1;
    } else if (is_unary(op)) {
      if (!need(a)) return fail("unary operand undefined");
      defined[dst] = 1;
    } else if (is_binary(op)) {
      if (!need(a) || !need(b)) return fail("binary operand undefined");
      defined[dst] = 1;
    } else {
      return fail("unknown opcode " + std::to_string(op));
    }
  }
  return true;
}

bool copy_sparsity(const bo_sparsity& in, int32_t n_rows, int32_t n_cols, bool lower_only, Sparsity* out,
                   std::string* err) {
  if (in.nnz < 0 || (in.nnz && (!in.row || !in.col))) {
    *err = "sparsity: bad arrays";
    return false;
  }
  out->row.assign(in.row, in.row + in.nnz);
  out->col.assign(in.col, in.col + in.nnz);
  for (int32_t k = 0; k < in.nnz; ++k) {
    if (out->row[k] < 0 || out->row[k] >= n_rows || out->col[k] < 0 || out->col[k] >= n_cols ||
        (lower_only && out->col[k] > out->row[k])) {
      *err = "sparsity: entry out of range";
      return false;
    }
  }
  return true;
}

static std::string fmt_const(double v) {
  if (std::isnan(v)) return "BO_NAN";
  if (std::isinf(v)) return v > 0 ? "BO_INF" : "(-BO_INF)";
  char buf[64];
  std::snprintf(buf, sizeof buf, "%.17g", v);
  std::string s(buf);
  if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
  if (v < 0) s = "(" + s + ")";
  return s;
}

TapeStats tape_stats(const Tape& tape) {
  TapeStats st;
  for (int64_t i = 0; i < tape.n_instr(); ++i) {
    const int op = tape.instr[4 * i] & 0xFF;
    switch (op) {
      case BO_OP_INPUT: case BO_OP_CONST: case BO_OP_OUTPUT: break;
      case BO_OP_ADD: case BO_OP_SUB: case BO_OP_MUL: case BO_OP_NEG: case BO_OP_SQ: st.n_arith++; break;
      case BO_OP_DIV: case BO_OP_SQRT: st.n_div_sqrt++; break;
      case BO_OP_SIN: case BO_OP_COS: case BO_OP_TAN: case BO_OP_ASIN: case BO_OP_ACOS: case BO_OP_ATAN:
      case BO_OP_ATAN2: case BO_OP_EXP: case BO_OP_LOG: case BO_OP_POW: case BO_OP_TANH: case BO_OP_SINH:
      case BO_OP_COSH: st.n_trig++; break;
      default: st.n_other++;
    }
  }
  return st;
}

std::string emit_tape_function(const Tape& tape, const std::string& name) { return emit_tape_function(tape, name, nullptr); }

std::string emit_tape_function(const Tape& tape, const std::string& name, const TapeEmitHooks* hooks) {
  const int64_t n = tape.n_instr();
  // pass 1: SSA operands (value ids = defining instruction index)
  std::vector<int64_t> cur(tape.n_work, -1), sa(n, -1), sb(n, -1), sc(n, -1);
  for (int64_t i = 0; i < n; ++i) {
    const int32_t* r = &tape.instr[4 * i];
    const int op = r[0] & 0xFF, c = (int)((uint32_t)r[0] >> 8);
    if (op == BO_OP_OUTPUT) {
      sa[i] = cur[r[1]];
      continue;
    }
    if (op == BO_OP_IF_ELSE) {
      sa[i] = cur[r[2]];
      sb[i] = cur[r[3]];
      sc[i] = cur[c];
    } else if (is_unary(op)) {
      sa[i] = cur[r[2]];
    } else if (is_binary(op)) {
      sa[i] = cur[r[2]];
      sb[i] = cur[r[3]];
    }
    cur[r[1]] = i;
  }
  // pass 2: pair sin/cos of the same SSA operand
  std::map<int64_t, std::pair<int64_t, int64_t>> trig;  // operand -> (sin instr, cos instr)
  for (int64_t i = 0; i < n; ++i) {
    const int op = tape.instr[4 * i] & 0xFF;
    if (op != BO_OP_SIN && op != BO_OP_COS) continue;
    auto& e = trig.emplace(sa[i], std::make_pair((int64_t)-1, (int64_t)-1)).first->second;
    int64_t& slot = (op == BO_OP_SIN) ? e.first : e.second;
    if (slot < 0) slot = i;
  }
  std::vector<char> emitted(n, 0);

  std::ostringstream o;
  if (hooks) {
    o << hooks->signature << " {\n" << hooks->prologue;
  } else {
    o << "BO_DEVICE void " << name << "(";
    bool first = true;
    for (size_t k = 0; k < tape.in_sizes.size(); ++k) {
      o << (first ? "" : ", ") << "const double* BO_RESTRICT i" << k;
      first = false;
    }
    for (size_t k = 0; k < tape.out_sizes.size(); ++k) {
      o << (first ? "" : ", ") << "double* BO_RESTRICT o" << k;
      first = false;
    }
    o << ") {\n";
  }
  auto V = [](int64_t id) { return "v" + std::to_string(id); };
  for (int64_t i = 0; i < n; ++i) {
    const int32_t* r = &tape.instr[4 * i];
    const int op = r[0] & 0xFF;
    if (emitted[i]) continue;
    const std::string a = sa[i] >= 0 ? V(sa[i]) : "", b = sb[i] >= 0 ? V(sb[i]) : "", c = sc[i] >= 0 ? V(sc[i]) : "";
    if (op == BO_OP_OUTPUT) {
      if (hooks) o << "  " << hooks->output_stmt(i, r, a) << "\n";
      else o << "  o" << r[3] << "[" << r[2] << "] = " << a << ";\n";
      continue;
    }
    if (op == BO_OP_SIN || op == BO_OP_COS) {
      const auto& e = trig[sa[i]];
      const int64_t partner = (op == BO_OP_SIN) ? e.second : e.first;
      const int64_t self_first = (op == BO_OP_SIN) ? e.first : e.second;
      if (self_first == i && partner > i) {
        const int64_t si = (op == BO_OP_SIN) ? i : partner, ci = (op == BO_OP_SIN) ? partner : i;
        o << "  double " << V(si) << ", " << V(ci) << "; bo_sincos(" << a << ", &" << V(si) << ", &" << V(ci) << ");\n";
        emitted[partner] = 1;
        continue;
      }
    }
    o << "  const double " << V(i) << " = ";
    switch (op) {
      case BO_OP_INPUT:
        if (hooks) o << hooks->input_expr(i, r);
        else o << "i" << r[3] << "[" << r[2] << "]";
        break;
      case BO_OP_CONST:
        if (hooks && !hooks->const_expr(i, r).empty()) o << hooks->const_expr(i, r);
        else o << fmt_const(tape.consts[r[2]]);
        break;
      case BO_OP_ADD: o << a << " + " << b; break;
      case BO_OP_SUB: o << a << " - " << b; break;
      case BO_OP_MUL: o << a << " * " << b; break;
      case BO_OP_DIV: o << a << " / " << b; break;
      case BO_OP_ATAN2: o << "atan2(" << a << ", " << b << ")"; break;
      case BO_OP_FMIN: o << "fmin(" << a << ", " << b << ")"; break;
      case BO_OP_FMAX: o << "fmax(" << a << ", " << b << ")"; break;
      case BO_OP_POW: o << "pow(" << a << ", " << b << ")"; break;
      case BO_OP_LT: o << "(double)(" << a << " < " << b << ")"; break;
      case BO_OP_LE: o << "(double)(" << a << " <= " << b << ")"; break;
      case BO_OP_EQ: o << "(double)(" << a << " == " << b << ")"; break;
      case BO_OP_NE: o << "(double)(" << a << " != " << b << ")"; break;
      case BO_OP_AND: o << "(double)((" << a << " != 0.0) && (" << b << " != 0.0))"; break;
      case BO_OP_OR: o << "(double)((" << a << " != 0.0) || (" << b << " != 0.0))"; break;
      case BO_OP_NEG: o << "-" << a; break;
      case BO_OP_SQ: o << a << " * " << a; break;
      case BO_OP_SQRT: o << "sqrt(" << a << ")"; break;
      case BO_OP_SIN: o << "sin(" << a << ")"; break;
      case BO_OP_COS: o << "cos(" << a << ")"; break;
      case BO_OP_TAN: o << "tan(" << a << ")"; break;
      case BO_OP_ASIN: o << "asin(" << a << ")"; break;
      case BO_OP_ACOS: o << "acos(" << a << ")"; break;
      case BO_OP_ATAN: o << "atan(" << a << ")"; break;
      case BO_OP_FABS: o << "fabs(" << a << ")"; break;
      case BO_OP_EXP: o << "exp(" << a << ")"; break;
      case BO_OP_LOG: o << "log(" << a << ")"; break;
      case BO_OP_NOT: o << "(double)(" << a << " == 0.0)"; break;
      case BO_OP_SIGN: o << "bo_sign(" << a << ")"; break;
      case BO_OP_FLOOR: o << "floor(" << a << ")"; break;
      case BO_OP_CEIL: o << "ceil(" << a << ")"; break;
      case BO_OP_TANH: o << "tanh(" << a << ")"; break;
      case BO_OP_SINH: o << "sinh(" << a << ")"; break;
      case BO_OP_COSH: o << "cosh(" << a << ")"; break;
      case BO_OP_IF_ELSE: o << "(" << c << " != 0.0 ? " << a << " : " << b << ")"; break;
      default: o << "BO_NAN /* bad op */";
    }
    o << ";\n";
  }
  o << "}\n";
  return o.str();
}

static void emit_sparse_helpers(std::ostringstream& o, const char* tag, const Sparsity& sp, int n_rows) {
  // out[col] += J[k] * v[row]
  o << "BO_DEVICE void bo_J" << tag << "t_acc(const double* BO_RESTRICT J, const double* BO_RESTRICT v, double* BO_RESTRICT out) {\n";
  for (int k = 0; k < sp.nnz(); ++k) o << "  out[" << sp.col[k] << "] += J[" << k << "] * v[" << sp.row[k] << "];\n";
  o << "  (void)J; (void)v; (void)out;\n}\n";
  // out[row] = sum J[k] * x[col]
  o << "BO_DEVICE void bo_J" << tag << "_mul(const double* BO_RESTRICT J, const double* BO_RESTRICT x, double* BO_RESTRICT out) {\n";
  std::vector<char> seen(n_rows > 0 ? n_rows : 1, 0);
  for (int k = 0; k < sp.nnz(); ++k) {
    o << "  out[" << sp.row[k] << "] " << (seen[sp.row[k]] ? "+=" : "=") << " J[" << k << "] * x[" << sp.col[k] << "];\n";
    seen[sp.row[k]] = 1;
  }
  for (int r = 0; r < n_rows; ++r)
    if (!seen[r]) o << "  out[" << r << "] = 0.0;\n";
  o << "  (void)J; (void)x; (void)out;\n}\n";
}

bool hessian_depends_on_eq_multipliers(const ProblemSource& ps) {
  for (int64_t i = 0; i < ps.kkt.n_instr(); ++i) {
    const int32_t* r = &ps.kkt.instr[4 * i];
    if ((r[0] & 0xFF) == BO_OP_INPUT && r[3] == 2) return true;  // input segment 2 of the kkt tape = y
  }
  return false;
}

bool problem_is_qp(const ProblemSource& ps) {
  // forward taint pass over the (slot-reusing) tape: bit k = "depends on input segment k" (0 x, 1 p, 2 y, 3 z)
  std::vector<uint8_t> taint((size_t)std::max(ps.kkt.n_work, 1), 0);
  for (int64_t i = 0; i < ps.kkt.n_instr(); ++i) {
    const int32_t* r = &ps.kkt.instr[4 * i];
    const int op = r[0] & 0xFF;
    if (op == BO_OP_INPUT) {
      taint[r[1]] = (uint8_t)(1u << r[3]);
    } else if (op == BO_OP_CONST) {
      taint[r[1]] = 0;
    } else if (op == BO_OP_OUTPUT) {
      if (r[3] >= 4 && (taint[r[1]] & 0x0D)) return false;  // output segments 4, 5, 6 = jac_eq, jac_ineq, hess
    } else if (op == BO_OP_IF_ELSE) {
      taint[r[1]] = taint[r[2]] | taint[r[3]] | taint[(uint32_t)r[0] >> 8];
    } else if (is_unary(op)) {
      taint[r[1]] = taint[r[2]];
    } else {
      taint[r[1]] = taint[r[2]] | taint[r[3]];
    }
  }
  return true;
}

std::string emit_problem_source(const ProblemSource& ps, int tpb, bool pivoted_ldl, const SparsePlan* sparse, bool large, bool qp) {
  std::ostringstream o;
  o << "// generated by libb200optas (bo_codegen.cpp): tier-S solver, one instance per thread\n";
  if (qp) o << "#define BO_QP 1\n";
  o << "#define BO_NX " << ps.nx << "\n#define BO_NP " << ps.np << "\n#define BO_ME " << ps.n_eq << "\n#define BO_MI "
    << ps.n_ineq << "\n#define BO_NNZ_JE " << ps.jac_eq.nnz() << "\n#define BO_NNZ_JI " << ps.jac_ineq.nnz()
    << "\n#define BO_NNZ_H " << ps.hess.nnz() << "\n#define BO_TPB " << tpb << "\n";
  if (pivoted_ldl) o << "#define BO_USE_BK 1\n";
  if (!hessian_depends_on_eq_multipliers(ps)) o << "#define BO_RECALC_DC_ONLY 1\n";
  if (sparse) o << "#define BO_SPARSE_LDL 1\n#define BO_SPARSE_VALS " << sparse->vals_size() << "\n";
  if (large) {
    // table-driven tier: no generated code at all, only the sizes
    o << "#define BO_LARGE 1\n#define BO_NWORK " << std::max(ps.fc.n_work, ps.kkt.n_work) << "\n";
    o << "#include \"bo_common.cuh\"\n#include \"bo_ipm_reg.cuh\"\n";
    return o.str();
  }
  o << "#include \"bo_common.cuh\"\n\n";
  o << emit_tape_function(ps.fc, "bo_tape_fc") << "\n";
  o << emit_tape_function(ps.kkt, "bo_tape_kkt") << "\n";
  emit_sparse_helpers(o, "E", ps.jac_eq, ps.n_eq);
  emit_sparse_helpers(o, "I", ps.jac_ineq, ps.n_ineq);

  const int nk = ps.nx + ps.n_eq;
  // index of the (i, j) entry, i >= j, in the per-instance matrix storage: packed lower triangle, or the
  // position in the sparse factor's value array
  auto kidx = [sparse](int i, int j) { return sparse ? sparse->pos(i, j) : i * (i + 1) / 2 + j; };
  const int ksize = sparse ? sparse->vals_size() : nk * (nk + 1) / 2;
  o << "BO_DEVICE void bo_kkt_fill(const double* BO_RESTRICT H, const double* BO_RESTRICT JE, const double* BO_RESTRICT JI,\n"
       "                           const double* BO_RESTRICT sigma, double* BO_RESTRICT K) {\n";
  o << (sparse ? "  BO_NOUNROLL\n" : "  BO_UNROLL\n") << "  for (int i = 0; i < " << ksize << "; ++i) K[i] = 0.0;\n";
  for (int k = 0; k < ps.hess.nnz(); ++k) o << "  K[" << kidx(ps.hess.row[k], ps.hess.col[k]) << "] += H[" << k << "];\n";
  {
    std::vector<std::vector<int>> by_row(ps.n_ineq > 0 ? ps.n_ineq : 1);
    for (int k = 0; k < ps.jac_ineq.nnz(); ++k) by_row[ps.jac_ineq.row[k]].push_back(k);
    for (int r = 0; r < ps.n_ineq; ++r) {
      const auto& ks = by_row[r];
      for (size_t u = 0; u < ks.size(); ++u)
        for (size_t w = 0; w < ks.size(); ++w) {
          const int cu = ps.jac_ineq.col[ks[u]], cw = ps.jac_ineq.col[ks[w]];
          if (cu < cw || (cu == cw && u < w)) continue;  // lower triangle, each unordered pair once
          if (cu == cw && u != w) continue;               // duplicate coordinates are not expected
          o << "  K[" << kidx(cu, cw) << "] += sigma[" << r << "] * JI[" << ks[u] << "] * JI[" << ks[w] << "];\n";
        }
    }
  }
  for (int k = 0; k < ps.jac_eq.nnz(); ++k)
    o << "  K[" << kidx(ps.nx + ps.jac_eq.row[k], ps.jac_eq.col[k]) << "] += JE[" << k << "];\n";
  o << "  (void)H; (void)JE; (void)JI; (void)sigma;\n}\n";

  // K[x,x] += rho * JE' JE (augmented-Lagrangian convexification of the (1,1) block, see bo_ipm_reg.cuh)
  o << "BO_DEVICE void bo_JEtJE_acc(const double* BO_RESTRICT JE, double rho, double* BO_RESTRICT K) {\n";
  {
    std::vector<std::vector<int>> by_row(ps.n_eq > 0 ? ps.n_eq : 1);
    for (int k = 0; k < ps.jac_eq.nnz(); ++k) by_row[ps.jac_eq.row[k]].push_back(k);
    for (int r = 0; r < ps.n_eq; ++r) {
      const auto& ks = by_row[r];
      for (size_t u = 0; u < ks.size(); ++u)
        for (size_t w = 0; w < ks.size(); ++w) {
          const int cu = ps.jac_eq.col[ks[u]], cw = ps.jac_eq.col[ks[w]];
          if (cu < cw || (cu == cw && u != w)) continue;
          o << "  K[" << kidx(cu, cw) << "] += rho * JE[" << ks[u] << "] * JE[" << ks[w] << "];\n";
        }
    }
  }
  o << "  (void)JE; (void)rho; (void)K;\n}\n";
  // out = H v (H: lower triangle of a symmetric matrix) -- the QP path moves the gradient along a step with it
  o << "BO_DEVICE void bo_H_mul(const double* BO_RESTRICT H, const double* BO_RESTRICT v, double* BO_RESTRICT out) {\n";
  o << "  for (int i = 0; i < " << ps.nx << "; ++i) out[i] = 0.0;\n";
  for (int k = 0; k < ps.hess.nnz(); ++k) {
    o << "  out[" << ps.hess.row[k] << "] += H[" << k << "] * v[" << ps.hess.col[k] << "];\n";
    if (ps.hess.row[k] != ps.hess.col[k]) o << "  out[" << ps.hess.col[k] << "] += H[" << k << "] * v[" << ps.hess.row[k] << "];\n";
  }
  o << "  (void)H; (void)v; (void)out;\n}\n";
  o << "\n";
  o << "#include \"bo_ipm_reg.cuh\"\n";
  return o.str();
}

std::string emit_function_source(const Tape& tape, int tpb, int out_stages) {
  std::ostringstream o;
  o << "// generated by libb200optas (bo_codegen.cpp): streaming evaluation kernel\n";
  int tot_in = 0, tot_out = 0;
  for (int s : tape.in_sizes) tot_in += s;
  for (int s : tape.out_sizes) tot_out += s;
  o << "#define BO_NIN " << tape.in_sizes.size() << "\n#define BO_NOUT " << tape.out_sizes.size() << "\n#define BO_TPB "
    << tpb << "\n#define BO_IN_TOTAL " << tot_in << "\n#define BO_OUT_TOTAL " << tot_out << "\n#define BO_OUT_STAGES "
    << out_stages << "\n";
  o << "#include \"bo_common.cuh\"\n";
  auto arr = [&](const char* nm, const std::vector<int32_t>& v) {
    o << "__device__ constexpr int " << nm << "[" << (v.empty() ? 1 : v.size()) << "] = {";
    for (size_t k = 0; k < v.size(); ++k) o << (k ? ", " : "") << v[k];
    if (v.empty()) o << "0";
    o << "};\n";
  };
  arr("BO_IN_SIZE", tape.in_sizes);
  arr("BO_OUT_SIZE", tape.out_sizes);
  o << "\n" << emit_tape_function(tape, "bo_tape_fn") << "\n";
  // per-thread call: segment k of thread t lives at base[k] + t * size[k] in shared memory
  o << "BO_DEVICE void bo_tape_call(double* const* si, double* const* so, int t) {\n  bo_tape_fn(";
  bool first = true;
  for (size_t k = 0; k < tape.in_sizes.size(); ++k) {
    o << (first ? "" : ", ") << "si[" << k << "] + t * " << tape.in_sizes[k];
    first = false;
  }
  for (size_t k = 0; k < tape.out_sizes.size(); ++k) {
    o << (first ? "" : ", ") << "so[" << k << "] + t * " << tape.out_sizes[k];
    first = false;
  }
  o << ");\n}\n\n#include \"bo_stream_eval.cuh\"\n";
  return o.str();
}

}  // namespace bo
