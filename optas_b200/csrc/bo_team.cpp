// bo_team.cpp -- code generation for the team tier (csrc/jit/bo_ipm_team.cuh): one problem instance per team of G
// threads that sit in G different warps, per-instance state in shared memory.
//
// The expression tapes (what CasADi's SX virtual machine interprets inside nlpsol on the reference path,
// optas/solver.py:395) are cut into G SLICES BY OUTPUT: slice r is the sub-tape that computes the outputs assigned to
// role r -- every instruction an output of the slice depends on, in the original order.  Sub-expressions shared by
// outputs of different slices (the forward-kinematics chain under all Jacobian / Hessian entries) are recomputed in
// each; a greedy longest-first assignment keeps the slices balanced and the duplication small.  Each slice becomes
// one straight-line function that reads its inputs from and writes its outputs to the team's shared-memory state.
#include "bo_team.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <sstream>

#include "bo_opcodes.h"

namespace bo {

namespace {

std::string fmt_double(double v) {
  char buf[64];
  std::snprintf(buf, sizeof buf, "%.17g", v);
  std::string t(buf);
  if (t.find_first_of(".eEn") == std::string::npos) t += ".0";
  return t;
}

bool is_unary(int op) { return op >= BO_OP_NEG && op <= BO_OP_COSH; }
bool is_binary(int op) { return op >= BO_OP_ADD && op <= BO_OP_OR; }

// rough issue-slot cost of one tape instruction on the FP64 pipe (FMA = 1)
int op_cost(int op) {
  switch (op) {
    case BO_OP_INPUT: return 1;  // a shared-memory load
    case BO_OP_CONST: case BO_OP_OUTPUT: return 0;
    case BO_OP_DIV: return 10;
    case BO_OP_SQRT: return 12;
    case BO_OP_SIN: case BO_OP_COS: return 25;  // fused into one sincos per operand
    case BO_OP_TAN: case BO_OP_ASIN: case BO_OP_ACOS: case BO_OP_ATAN: case BO_OP_ATAN2: case BO_OP_EXP: case BO_OP_LOG:
    case BO_OP_POW: case BO_OP_TANH: case BO_OP_SINH: case BO_OP_COSH: return 50;
    default: return 1;
  }
}

struct Bits {
  std::vector<uint64_t> w;
  explicit Bits(size_t n = 0) : w((n + 63) / 64, 0) {}
  void set(size_t i) { w[i >> 6] |= 1ULL << (i & 63); }
  bool get(size_t i) const { return (w[i >> 6] >> (i & 63)) & 1ULL; }
};

}  // namespace

TeamSlices slice_tape(const Tape& tape, int G, const std::vector<int>& extra_cost_per_output_segment) {
  const int64_t n = tape.n_instr();
  TeamSlices ts;
  ts.G = G;
  // SSA operands (value id = defining instruction)
  std::vector<int64_t> cur(tape.n_work, -1), sa(n, -1), sb(n, -1), sc(n, -1);
  std::vector<int> cost(n, 0);
  std::vector<int64_t> outputs;
  for (int64_t i = 0; i < n; ++i) {
    const int32_t* r = &tape.instr[4 * i];
    const int op = r[0] & 0xFF, c = (int)((uint32_t)r[0] >> 8);
    cost[i] = op_cost(op);
    if (op == BO_OP_OUTPUT) {
      sa[i] = cur[r[1]];
      outputs.push_back(i);
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
  // ancestor set of every output (the defining instruction of an operand always precedes its user)
  std::vector<Bits> anc(outputs.size(), Bits((size_t)n));
  std::vector<int64_t> stack;
  std::vector<int> own_cost(outputs.size(), 0);
  for (size_t o = 0; o < outputs.size(); ++o) {
    Bits& b = anc[o];
    stack.clear();
    stack.push_back(outputs[o]);
    while (!stack.empty()) {
      const int64_t i = stack.back();
      stack.pop_back();
      if (i < 0 || b.get((size_t)i)) continue;
      b.set((size_t)i);
      own_cost[o] += cost[i];
      if (sa[i] >= 0) stack.push_back(sa[i]);
      if (sb[i] >= 0) stack.push_back(sb[i]);
      if (sc[i] >= 0) stack.push_back(sc[i]);
    }
    const int seg = tape.instr[4 * outputs[o] + 3];
    if (seg >= 0 && seg < (int)extra_cost_per_output_segment.size()) own_cost[o] += extra_cost_per_output_segment[seg];
  }
  // greedy: most expensive output first, to the role whose load (cost of the union of ancestors + extras) grows least
  std::vector<size_t> order(outputs.size());
  for (size_t o = 0; o < order.size(); ++o) order[o] = o;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return own_cost[a] > own_cost[b]; });
  std::vector<Bits> have(G, Bits((size_t)n));
  std::vector<int64_t> load(G, 0);
  ts.owner.assign(outputs.size(), 0);
  for (size_t o : order) {
    int best = 0;
    int64_t best_load = -1, best_add = 0;
    const int seg = tape.instr[4 * outputs[o] + 3];
    const int extra = (seg >= 0 && seg < (int)extra_cost_per_output_segment.size()) ? extra_cost_per_output_segment[seg] : 0;
    for (int r = 0; r < G; ++r) {
      int64_t add = extra;
      for (size_t w = 0; w < anc[o].w.size(); ++w) {
        uint64_t fresh = anc[o].w[w] & ~have[r].w[w];
        while (fresh) {
          const int bit = __builtin_ctzll(fresh);
          add += cost[w * 64 + bit];
          fresh &= fresh - 1;
        }
      }
      const int64_t nl = load[r] + add;
      if (best_load < 0 || nl < best_load) {
        best_load = nl;
        best = r;
        best_add = add;
      }
    }
    ts.owner[o] = best;
    load[best] += best_add;
    for (size_t w = 0; w < anc[o].w.size(); ++w) have[best].w[w] |= anc[o].w[w];
  }
  // sub-tapes: the instructions of the union, original order
  ts.slices.resize(G);
  ts.cost = load;
  ts.total_cost = 0;
  for (int64_t i = 0; i < n; ++i) ts.total_cost += cost[i];
  for (int r = 0; r < G; ++r) {
    Tape& t = ts.slices[r];
    t.consts = tape.consts;
    t.n_work = tape.n_work;
    t.in_sizes = tape.in_sizes;
    t.out_sizes = tape.out_sizes;
    for (int64_t i = 0; i < n; ++i)
      if (have[r].get((size_t)i)) t.instr.insert(t.instr.end(), &tape.instr[4 * i], &tape.instr[4 * i] + 4);
  }
  ts.output_instr = outputs;
  for (int64_t i : outputs) {
    ts.out_seg.push_back(tape.instr[4 * i + 3]);
    ts.out_elem.push_back(tape.instr[4 * i + 2]);
  }
  return ts;
}

// ---------------------------------------------------------------------------------------------------------------
// tape surgery: SSA form, shared sin/cos of the decision variables, KKT (1,1) block as tape outputs
// ---------------------------------------------------------------------------------------------------------------
namespace {

// A tape under construction in pure SSA form (work slot = row index).
struct Builder {
  std::vector<int32_t> rows;  // [n][4]
  std::vector<double> consts;
  int add(int op, int a = 0, int b = 0, int c = 0) {
    const int id = (int)(rows.size() / 4);
    rows.push_back(op | (c << 8));
    rows.push_back(id);
    rows.push_back(a);
    rows.push_back(b);
    return id;
  }
  int constant(double v) {
    consts.push_back(v);
    return add(BO_OP_CONST, (int)consts.size() - 1, 0);
  }
  void output(int value, int elem, int seg) {
    rows.push_back(BO_OP_OUTPUT);
    rows.push_back(value);
    rows.push_back(elem);
    rows.push_back(seg);
  }
};

// Re-emit `t` in SSA form into `b`; returns, per OUTPUT row of t in order, (segment, element, value id) and lets the
// caller decide which outputs to keep.  sin / cos whose operand is directly element k of input segment 0 (x) are
// replaced by reads of a new input segment `trig_seg` (element 2k = sin, 2k + 1 = cos) and k is added to *trig_mask.
struct OutRef { int seg, elem, value; };
std::vector<OutRef> to_ssa(const Tape& t, Builder* b, int trig_seg, uint32_t* trig_mask) {
  std::vector<int> cur(t.n_work, -1);
  std::vector<int> x_elem_of;  // value id -> element of x it reads directly, or -1
  std::vector<OutRef> outs;
  b->consts = t.consts;
  auto note = [&](int id, int elem) {
    if ((int)x_elem_of.size() <= id) x_elem_of.resize(id + 1, -1);
    x_elem_of[id] = elem;
  };
  for (int64_t i = 0; i < t.n_instr(); ++i) {
    const int32_t* r = &t.instr[4 * i];
    const int op = r[0] & 0xFF, c = (int)((uint32_t)r[0] >> 8);
    if (op == BO_OP_OUTPUT) {
      outs.push_back({r[3], r[2], cur[r[1]]});
      continue;
    }
    int id;
    if (op == BO_OP_INPUT) {
      id = b->add(BO_OP_INPUT, r[2], r[3]);
      note(id, r[3] == 0 ? r[2] : -1);
    } else if (op == BO_OP_CONST) {
      id = b->add(BO_OP_CONST, r[2], 0);
      note(id, -1);
    } else if ((op == BO_OP_SIN || op == BO_OP_COS) && trig_seg >= 0 && x_elem_of[cur[r[2]]] >= 0 && x_elem_of[cur[r[2]]] < 16) {
      const int k = x_elem_of[cur[r[2]]];
      id = b->add(BO_OP_INPUT, 2 * k + (op == BO_OP_COS ? 1 : 0), trig_seg);
      *trig_mask |= 1u << k;
      note(id, -1);
    } else if (op == BO_OP_IF_ELSE) {
      id = b->add(op, cur[r[2]], cur[r[3]], cur[c]);
      note(id, -1);
    } else if (is_unary(op)) {
      id = b->add(op, cur[r[2]], 0);
      note(id, -1);
    } else {
      id = b->add(op, cur[r[2]], cur[r[3]]);
      note(id, -1);
    }
    cur[r[1]] = id;
  }
  return outs;
}

Tape finish(const Builder& b, std::vector<int32_t> in_sizes, std::vector<int32_t> out_sizes) {
  Tape t;
  t.instr = b.rows;
  t.consts = b.consts;
  t.n_work = (int32_t)(b.rows.size() / 4) + 1;
  t.in_sizes = std::move(in_sizes);
  t.out_sizes = std::move(out_sizes);
  return t;
}

}  // namespace

// kkt tape of the team tier: inputs (x, p, y, z, sigma, trig), outputs (f, grad, cE, cI, JE nz, JI nz, KX) with
//   KX = packed lower triangle of  H + JI' diag(sigma) JI + rho JE'JE   -- the (1,1) block of the rho-augmented KKT matrix,
// so that its assembly is sliced over the roles together with the derivatives it is made of.
static Tape team_kkt_tape(const ProblemSource& ps, double rho, uint32_t* trig_mask) {
  Builder b;
  const std::vector<OutRef> outs = to_ssa(ps.kkt, &b, 5, trig_mask);
  std::vector<int> je(ps.jac_eq.nnz(), -1), ji(ps.jac_ineq.nnz(), -1), hh(ps.hess.nnz(), -1);
  for (const OutRef& o : outs) {
    if (o.seg <= 5) b.output(o.value, o.elem, o.seg);
    if (o.seg == 4) je[o.elem] = o.value;
    if (o.seg == 5) ji[o.elem] = o.value;
    if (o.seg == 6) hh[o.elem] = o.value;
  }
  const int nx = ps.nx;
  std::vector<int> acc((size_t)nx * (nx + 1) / 2, -1);
  auto kidx = [](int i, int j) { return i * (i + 1) / 2 + j; };
  auto add_term = [&](int i, int j, int v) {
    int& a = acc[kidx(i, j)];
    a = a < 0 ? v : b.add(BO_OP_ADD, a, v);
  };
  for (int k = 0; k < ps.hess.nnz(); ++k) add_term(ps.hess.row[k], ps.hess.col[k], hh[k]);
  {
    std::vector<std::vector<int>> by_row(ps.n_ineq > 0 ? ps.n_ineq : 1);
    for (int k = 0; k < ps.jac_ineq.nnz(); ++k) by_row[ps.jac_ineq.row[k]].push_back(k);
    for (int r = 0; r < ps.n_ineq; ++r) {
      const auto& ks = by_row[r];
      if (ks.empty()) continue;
      const int sig = b.add(BO_OP_INPUT, r, 4);
      for (size_t u = 0; u < ks.size(); ++u)
        for (size_t w = 0; w < ks.size(); ++w) {
          const int cu = ps.jac_ineq.col[ks[u]], cw = ps.jac_ineq.col[ks[w]];
          if (cu < cw || (cu == cw && u != w)) continue;
          add_term(cu, cw, b.add(BO_OP_MUL, sig, b.add(BO_OP_MUL, ji[ks[u]], ji[ks[w]])));
        }
    }
  }
  if (ps.n_eq > 0) {
    const int rho_id = b.constant(rho);
    std::vector<std::vector<int>> by_row(ps.n_eq);
    for (int k = 0; k < ps.jac_eq.nnz(); ++k) by_row[ps.jac_eq.row[k]].push_back(k);
    for (int r = 0; r < ps.n_eq; ++r) {
      const auto& ks = by_row[r];
      for (size_t u = 0; u < ks.size(); ++u)
        for (size_t w = 0; w < ks.size(); ++w) {
          const int cu = ps.jac_eq.col[ks[u]], cw = ps.jac_eq.col[ks[w]];
          if (cu < cw || (cu == cw && u != w)) continue;
          add_term(cu, cw, b.add(BO_OP_MUL, rho_id, b.add(BO_OP_MUL, je[ks[u]], je[ks[w]])));
        }
    }
  }
  int zero = -1;
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j <= i; ++j) {
      int v = acc[kidx(i, j)];
      if (v < 0) {
        if (zero < 0) zero = b.constant(0.0);
        v = zero;
      }
      b.output(v, kidx(i, j), 6);
    }
  return finish(b, {ps.nx, ps.np, ps.n_eq, ps.n_ineq, ps.n_ineq, 2 * ps.nx},
                {1, ps.nx, ps.n_eq, ps.n_ineq, ps.jac_eq.nnz(), ps.jac_ineq.nnz(), nx * (nx + 1) / 2});
}

static Tape team_fc_tape(const ProblemSource& ps, uint32_t* trig_mask) {
  Builder b;
  const std::vector<OutRef> outs = to_ssa(ps.fc, &b, 2, trig_mask);
  for (const OutRef& o : outs) b.output(o.value, o.elem, o.seg);
  return finish(b, {ps.nx, ps.np, 2 * ps.nx}, {1, ps.n_eq, ps.n_ineq});
}

namespace {

std::string at(const char* off, int k) { return std::string("SM(") + off + ", " + std::to_string(k) + ")"; }

// one slice -> `BO_NOINLINE void <name>(double* sm [, double at, bool rows])`
std::string emit_slice(const Tape& slice, const TeamSlices& ts, const std::string& name, bool kkt, int role) {
  TapeEmitHooks h;
  if (kkt) {
    h.signature = "BO_NOINLINE void " + name + "(double* BO_RESTRICT sm)";
  } else {
    h.signature = "BO_NOINLINE void " + name + "(double* BO_RESTRICT sm, const double at, const bool rows)";
  }
  h.prologue = "";
  h.input_expr = [kkt](int64_t, const int32_t* r) -> std::string {
    const int k = r[2], seg = r[3];
    if (kkt) {
      static const char* offs[6] = {"BO_OFF_X", "BO_OFF_P", "BO_OFF_Y", "BO_OFF_Z", "BO_OFF_SIG", "BO_OFF_SH"};
      return at(offs[seg], k);
    }
    static const char* offs[3] = {"BO_OFF_XT", "BO_OFF_P", "BO_OFF_SH"};
    return at(offs[seg], k);
  };
  // Double literals whose low word is not zero cost two moves each in SASS (UMOV + IMAD.MOV); read from the constant
  // bank they are a free operand of DFMA / DMUL / DADD.  Literals with a zero low word (0, 1, 0.5, 2 ...) are immediates.
  h.const_expr = [kkt, &slice](int64_t, const int32_t* r) -> std::string {
    const double v = slice.consts[r[2]];
    uint64_t bits;
    static_assert(sizeof bits == sizeof v, "");
    __builtin_memcpy(&bits, &v, sizeof bits);
    if ((bits & 0xFFFFFFFFULL) == 0ULL || !std::isfinite(v)) return "";
    return std::string(kkt ? "BO_CK[" : "BO_CF[") + std::to_string(r[2]) + "]";
  };
  h.output_stmt = [kkt](int64_t, const int32_t* r, const std::string& v) -> std::string {
    const int k = r[2], seg = r[3];
    if (kkt) {
      static const char* offs[7] = {"BO_OFF_F0", "BO_OFF_G", "BO_OFF_CE", "BO_OFF_CI", "BO_OFF_JE", "BO_OFF_JI", "BO_OFF_KX"};
      return at(offs[seg], k) + " = " + v + ";";
    }
    static const char* offs[3] = {"BO_OFF_FT", "BO_OFF_CET", "BO_OFF_CIT"};
    return at(offs[seg], k) + " = " + v + ";";
  };
  std::string body = emit_tape_function(slice, name, &h);
  // every slice ends with the per-row work of the constraint rows it owns (shared, non-inlined: bo_ipm_team.cuh)
  unsigned long long mask_e = 0, mask_i = 0;
  const int seg_e = kkt ? 2 : 1, seg_i = kkt ? 3 : 2;
  for (size_t o = 0; o < ts.owner.size(); ++o) {
    if (ts.owner[o] != role) continue;
    if (ts.out_seg[o] == seg_e) mask_e |= 1ULL << ts.out_elem[o];
    if (ts.out_seg[o] == seg_i) mask_i |= 1ULL << ts.out_elem[o];
  }
  const size_t close = body.rfind('}');
  std::ostringstream tail;
  if (kkt) tail << "  bo_team_rows_kkt(sm, " << role << ", 0x" << std::hex << mask_e << "ULL, 0x" << mask_i << std::dec << "ULL);\n}\n";
  else tail << "  bo_team_rows(sm, at, rows, " << role << ", 0x" << std::hex << mask_e << "ULL, 0x" << mask_i << std::dec << "ULL);\n}\n";
  return body.substr(0, close) + tail.str();
}

// strided sparse helpers: J lives in shared memory (stride BO_LS); v / out have their own strides
void emit_sparse_helpers_t(std::ostringstream& o, const char* tag, const Sparsity& sp, int n_rows) {
  o << "BO_DEVICE void bo_J" << tag << "t_acc_t(const double* BO_RESTRICT J, const double* BO_RESTRICT v, const int vs, const double sg, "
       "double* BO_RESTRICT out, const int os) {\n";
  for (int k = 0; k < sp.nnz(); ++k)
    o << "  out[" << sp.col[k] << " * os] += sg * (J[" << k << " * BO_LS] * v[" << sp.row[k] << " * vs]);\n";
  o << "  (void)J; (void)v; (void)out; (void)vs; (void)os; (void)sg;\n}\n";
  o << "BO_DEVICE void bo_J" << tag << "_mul_t(const double* BO_RESTRICT J, const double* BO_RESTRICT x, const int xs, "
       "double* BO_RESTRICT out, const int os) {\n";
  std::vector<std::vector<int>> by_row(n_rows > 0 ? n_rows : 1);
  for (int k = 0; k < sp.nnz(); ++k) by_row[sp.row[k]].push_back(k);
  for (int r = 0; r < n_rows; ++r) {
    o << "  out[" << r << " * os] = ";
    if (by_row[r].empty()) o << "0.0";
    for (size_t u = 0; u < by_row[r].size(); ++u)
      o << (u ? " + " : "") << "J[" << by_row[r][u] << " * BO_LS] * x[" << sp.col[by_row[r][u]] << " * xs]";
    o << ";\n";
  }
  o << "  (void)J; (void)x; (void)out; (void)xs; (void)os;\n}\n";
}

}  // namespace

std::string emit_team_source(const ProblemSource& ps, const TeamPlan& plan) {
  std::ostringstream o;
  const int G = plan.G;
  o << "// generated by libb200optas (bo_team.cpp): team tier, one instance per " << G << " threads in " << G << " warps\n";
  o << "#define BO_TEAM 1\n#define BO_NX " << ps.nx << "\n#define BO_NP " << ps.np << "\n#define BO_ME " << ps.n_eq << "\n#define BO_MI "
    << ps.n_ineq << "\n#define BO_NNZ_JE " << ps.jac_eq.nnz() << "\n#define BO_NNZ_JI " << ps.jac_ineq.nnz()
    << "\n#define BO_G " << G << "\n#define BO_TPB " << 32 * G << "\n#define BO_TRIG_MASK 0x" << std::hex << plan.trig_mask
    << std::dec << "u\n#define BO_STATIC_RHO " << fmt_double(plan.rho) << "\n";
  if (!hessian_depends_on_eq_multipliers(ps)) o << "#define BO_RECALC_DC_ONLY 1\n";
  o << "// slice costs (FP64 issue slots, est.): kkt";
  for (int r = 0; r < G; ++r) o << " " << plan.kkt.cost[r];
  o << " of " << plan.kkt.total_cost << " unsliced; fc";
  for (int r = 0; r < G; ++r) o << " " << plan.fc.cost[r];
  o << " of " << plan.fc.total_cost << "\n";
  o << "#include \"bo_common.cuh\"\n#include \"bo_team_layout.cuh\"\n";
  o << "static_assert(BO_SM_ELEMS == " << team_smem_elems(ps, G) << ", \"host / device shared-memory layouts disagree\");\n";
  o << "\n";
  auto pool = [&](const char* nm, const std::vector<double>& c) {
    o << "BO_CONSTANT double " << nm << "[" << std::max<size_t>(1, c.size()) << "] = {";
    for (size_t k = 0; k < c.size(); ++k) o << (k ? ", " : "") << (std::isfinite(c[k]) ? fmt_double(c[k]) : std::string("0.0"));
    if (c.empty()) o << "0.0";
    o << "};\n";
  };
  pool("BO_CK", plan.kkt_tape.consts);
  pool("BO_CF", plan.fc_tape.consts);
  o << "BO_NOINLINE void bo_team_rows(double* BO_RESTRICT sm, double at, bool rows, int role, unsigned long long mask_e, unsigned long long mask_i);\n";
  o << "BO_NOINLINE void bo_team_rows_kkt(double* BO_RESTRICT sm, int role, unsigned long long mask_e, unsigned long long mask_i);\n\n";
  for (int r = 0; r < G; ++r) o << emit_slice(plan.kkt.slices[r], plan.kkt, "bo_kkt_r" + std::to_string(r), true, r) << "\n";
  for (int r = 0; r < G; ++r) o << emit_slice(plan.fc.slices[r], plan.fc, "bo_fc_r" + std::to_string(r), false, r) << "\n";
  o << "BO_DEVICE void bo_team_kkt(const int role, double* BO_RESTRICT sm) {\n  switch (role) {\n";
  for (int r = 0; r < G; ++r) o << "    case " << r << ": bo_kkt_r" << r << "(sm); break;\n";
  o << "    default: break;\n  }\n}\n";
  o << "BO_DEVICE void bo_team_fc(const int role, double* BO_RESTRICT sm, const double at, const bool rows) {\n  switch (role) {\n";
  for (int r = 0; r < G; ++r) o << "    case " << r << ": bo_fc_r" << r << "(sm, at, rows); break;\n";
  o << "    default: break;\n  }\n}\n\n";
  emit_sparse_helpers_t(o, "E", ps.jac_eq, ps.n_eq);
  emit_sparse_helpers_t(o, "I", ps.jac_ineq, ps.n_ineq);
  // the (2,1) block of the KKT matrix: JE entries into the packed lower triangle held in registers
  o << "BO_DEVICE void bo_kkt_je_t(const double* BO_RESTRICT JE, double* BO_RESTRICT K) {\n";
  for (int k = 0; k < ps.jac_eq.nnz(); ++k) {
    const int i = ps.nx + ps.jac_eq.row[k], j = ps.jac_eq.col[k];
    o << "  K[" << i * (i + 1) / 2 + j << "] += JE[" << k << " * BO_LS];\n";
  }
  o << "  (void)JE; (void)K;\n}\n\n";
  // Gauss-Newton (1,1) block of the restoration phase: JI' diag(w) JI + rho JE'JE into the packed lower triangle
  // (K: strided shared memory, e.g. the KX array)
  o << "BO_NOINLINE void bo_kkt_gn_t(const double* BO_RESTRICT JE, const double* BO_RESTRICT JI, const double* BO_RESTRICT w, const double rho, double* BO_RESTRICT K) {\n";
  o << "  BO_NOUNROLL\n  for (int i = 0; i < " << ps.nx * (ps.nx + 1) / 2 << "; ++i) K[i * BO_LS] = 0.0;\n";
  {
    auto kidx = [](int i, int j) { return i * (i + 1) / 2 + j; };
    std::vector<std::vector<int>> by_row(ps.n_ineq > 0 ? ps.n_ineq : 1);
    for (int k = 0; k < ps.jac_ineq.nnz(); ++k) by_row[ps.jac_ineq.row[k]].push_back(k);
    for (int r = 0; r < ps.n_ineq; ++r) {
      const auto& ks = by_row[r];
      for (size_t u = 0; u < ks.size(); ++u)
        for (size_t v = 0; v < ks.size(); ++v) {
          const int cu = ps.jac_ineq.col[ks[u]], cv = ps.jac_ineq.col[ks[v]];
          if (cu < cv || (cu == cv && u != v)) continue;
          o << "  K[" << kidx(cu, cv) << " * BO_LS] += w[" << r << " * BO_LS] * (JI[" << ks[u] << " * BO_LS] * JI[" << ks[v] << " * BO_LS]);\n";
        }
    }
    std::vector<std::vector<int>> by_row_e(ps.n_eq > 0 ? ps.n_eq : 1);
    for (int k = 0; k < ps.jac_eq.nnz(); ++k) by_row_e[ps.jac_eq.row[k]].push_back(k);
    for (int r = 0; r < ps.n_eq; ++r) {
      const auto& ks = by_row_e[r];
      for (size_t u = 0; u < ks.size(); ++u)
        for (size_t v = 0; v < ks.size(); ++v) {
          const int cu = ps.jac_eq.col[ks[u]], cv = ps.jac_eq.col[ks[v]];
          if (cu < cv || (cu == cv && u != v)) continue;
          o << "  K[" << kidx(cu, cv) << " * BO_LS] += rho * (JE[" << ks[u] << " * BO_LS] * JE[" << ks[v] << " * BO_LS]);\n";
        }
    }
  }
  o << "  (void)JE; (void)JI; (void)w; (void)rho; (void)K;\n}\n\n";
  o << "#include \"bo_ipm_team.cuh\"\n";
  return o.str();
}

bool make_team_plan(const ProblemSource& ps, int G, TeamPlan* plan, std::string* why) {
  plan->G = G;
  plan->rho = 1.0e6;
  if (ps.n_ineq > 64 || ps.n_eq > 64 || ps.nx > 16) {
    *why = "more than 64 constraint rows of a kind or more than 16 variables";
    return false;
  }
  // every element of every output segment must be written by the tape (the slices own rows by their OUTPUT rows)
  auto covered = [&](const Tape& t, const char* nm) {
    std::vector<std::vector<char>> seen(t.out_sizes.size());
    for (size_t s = 0; s < seen.size(); ++s) seen[s].assign((size_t)t.out_sizes[s], 0);
    for (int64_t i = 0; i < t.n_instr(); ++i) {
      const int32_t* r = &t.instr[4 * i];
      if ((r[0] & 0xFF) == BO_OP_OUTPUT) seen[r[3]][r[2]] = 1;
    }
    for (size_t s = 0; s < seen.size(); ++s)
      for (char c : seen[s])
        if (!c) {
          *why = std::string(nm) + " tape leaves an output element unwritten";
          return false;
        }
    return true;
  };
  if (!covered(ps.kkt, "kkt") || !covered(ps.fc, "fc")) return false;
  plan->trig_mask = 0;
  plan->kkt_tape = team_kkt_tape(ps, plan->rho, &plan->trig_mask);
  plan->fc_tape = team_fc_tape(ps, &plan->trig_mask);
  plan->kkt = slice_tape(plan->kkt_tape, G, {0, 0, 4, 10});
  // per-row barrier work that rides with the owner of a row of the f / c tape: two logs and a division
  plan->fc = slice_tape(plan->fc_tape, G, {0, 2, 110});
  return true;
}

int team_smem_elems(const ProblemSource& ps, int G) {
  const int nk = ps.nx + ps.n_eq;
  const int ksz = nk * (nk + 1) / 2;
  // must match the BO_OFF_* chain of bo_team_layout.cuh
  return ps.np + 2 * ps.nx /* X XT */ + 2 * ps.nx /* SH */ + 4 * ps.n_ineq /* S Z RS SIG */ + ps.n_eq /* Y */ + 2 * ps.n_ineq /* SN RSN */ +
         1 + ps.nx + ps.n_eq + ps.n_ineq + ps.jac_eq.nnz() + ps.jac_ineq.nnz() + ps.nx * (ps.nx + 1) / 2 /* F0 G CE CI JE JI KX */ +
         ps.nx + ksz + nk /* RD LD SOL */ + ps.nx + ps.n_ineq + ps.n_eq + ps.nx + ps.n_ineq + ps.n_eq + ps.n_ineq /* DX DS YST DX0 DS0 RE RI */ +
         1 + ps.n_eq + ps.n_ineq /* FT CET CIT */ + 5 * G + 1 + 8 + 8 /* PART AT FTH FPH */;
}

}  // namespace bo
