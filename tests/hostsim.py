"""Host-compiled build of the GENERATED solver source -- a test harness, not a product path.

The tier-S solver template (optas_b200/csrc/jit/bo_ipm_reg.cuh) and the code generated from a
problem's tapes are plain C++ apart from the kernel entry point, so with ``-DBO_HOST_SIM`` the very
same text compiles with g++.  GPU-less CI uses this to exercise the solver *logic* (convergence,
inertia correction, line search) on the exact source that NVRTC compiles for sm_100a.  Nothing in
``optas_b200`` imports this module; ``B200Solver`` never runs on the CPU.
"""

from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_JIT_INC = os.path.join(_HERE, "..", "optas_b200", "csrc", "jit")

_WRAPPER = r"""
#define BO_HOST_SIM 1
#include <cstdio>
%(trace)s
#include "%(gen)s"
extern "C" void hostsim_solve(long long B, const double* p, const double* x0, double* x, double* lam, double* f,
                              int* status, int* iters, double* kkt, int* trips, int max_iter, double tol, double acc_tol,
                              double mu_init, double max_step, int max_trips, const int* ldl_tab, const double* dtab, double* scratch) {
  bo_solver_params prm;
  prm.max_iter = max_iter; prm.tol = tol; prm.acceptable_tol = acc_tol; prm.mu_init = mu_init; prm.max_step = max_step; prm.max_trips = max_trips; prm.ldl_tab = ldl_tab; prm.dtab = dtab; prm.scratch = scratch; prm.scratch_stride = 1;
  for (long long b = 0; b < B; ++b) {
    static bo_ipm_state S;
#ifdef BO_LARGE
    S.ldp = scratch; S.lds = 1;
#endif
    for (int i = 0; i < BO_NP; ++i) S.p[i] = p[b * BO_NP + i];
    for (int i = 0; i < BO_NX; ++i) S.x[i] = x0 ? x0[b * BO_NX + i] : 0.0;
    status[b] = bo_ipm_solve(S, prm);
    for (int i = 0; i < BO_NX; ++i) x[b * BO_NX + i] = S.x[i];
    for (int j = 0; j < BO_ME; ++j) lam[b * (BO_ME + BO_MI) + j] = S.y[j];
    for (int i = 0; i < BO_MI; ++i) lam[b * (BO_ME + BO_MI) + BO_ME + i] = S.z[i];
    f[b] = S.f; iters[b] = S.it; kkt[b] = S.err0; if (trips) trips[b] = S.trips;
  }
}
"""

_WRAPPER_TEAM = r"""
#define BO_HOST_SIM 1
#include <cstdio>
#include <vector>
%(trace)s
#include "%(gen)s"
// team tier: the G roles of a team run one after the other where the kernel has a CTA barrier (bo_team_solve_host)
extern "C" void hostsim_solve(long long B, const double* p, const double* x0, double* x, double* lam, double* f,
                              int* status, int* iters, double* kkt, int* trips, int max_iter, double tol, double acc_tol,
                              double mu_init, double max_step, int max_trips, const int* ldl_tab, const double* dtab, double* scratch) {
  bo_solver_params prm;
  prm.max_iter = max_iter; prm.tol = tol; prm.acceptable_tol = acc_tol; prm.mu_init = mu_init; prm.max_step = max_step; prm.max_trips = max_trips; prm.ldl_tab = ldl_tab; prm.dtab = dtab; prm.scratch = scratch; prm.scratch_stride = 1;
  std::vector<double> smv(BO_SM_ELEMS + 1);
  double* sm = smv.data();
  for (long long b = 0; b < B; ++b) {
    for (auto& v : smv) v = 0.0 / 0.0;  // anything read before it is written shows up as NaN
    bo_tm M;
    for (int i = 0; i < BO_NP; ++i) SM(BO_OFF_P, i) = p[b * BO_NP + i];
    for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_X, i) = x0 ? x0[b * BO_NX + i] : 0.0;
    status[b] = bo_team_solve_host(M, sm, prm);
    for (int i = 0; i < BO_NX; ++i) x[b * BO_NX + i] = SM(BO_OFF_X, i);
    for (int j = 0; j < BO_ME; ++j) lam[b * (BO_ME + BO_MI) + j] = SM(BO_OFF_Y, j);
    for (int i = 0; i < BO_MI; ++i) lam[b * (BO_ME + BO_MI) + BO_ME + i] = SM(BO_OFF_Z, i);
    f[b] = M.f; iters[b] = M.it; kkt[b] = M.err0; if (trips) trips[b] = M.trips;
  }
}
"""

_WRAPPER_COOP = r"""
#define BO_HOST_SIM 1
#include <cstdio>
#include <vector>
%(trace)s
#include "%(gen)s"
extern "C" void hostsim_solve(long long B, const double* p, const double* x0, double* x, double* lam, double* f,
                              int* status, int* iters, double* kkt, int* trips, int max_iter, double tol, double acc_tol,
                              double mu_init, double max_step, int max_trips, const int* ldl_tab, const double* dtab, double* scratch) {
  bo_solver_params prm;
  prm.max_iter = max_iter; prm.tol = tol; prm.acceptable_tol = acc_tol; prm.mu_init = mu_init; prm.max_step = max_step; prm.max_trips = max_trips; prm.ldl_tab = ldl_tab; prm.dtab = dtab; prm.scratch = scratch; prm.scratch_stride = 1;
  std::vector<double> W(BO_SCRATCH_DOUBLES), vals(BO_VALS + 1 + BO_NK + 1), red(64);  // bp right behind vals, as in shared memory
  int ibuf[4];
  bo_cta C;
  C.W = W.data(); C.vals = vals.data(); C.bp = vals.data() + BO_VALS + 1; C.red = red.data(); C.ibuf = ibuf; C.tab = ldl_tab; C.dtab = dtab; C.prof = nullptr; C.wkkt = nullptr; C.wfc = nullptr;
  for (long long b = 0; b < B; ++b) {
    bo_cta_state S;
    for (int i = 0; i < BO_NP; ++i) W[BO_OFF_P + i] = p[b * BO_NP + i];
    for (int i = 0; i < BO_NX; ++i) W[BO_OFF_X + i] = x0 ? x0[b * BO_NX + i] : 0.0;
    status[b] = bo_cta_solve(S, C, prm);
    for (int i = 0; i < BO_NX; ++i) x[b * BO_NX + i] = W[BO_OFF_X + i];
    for (int j = 0; j < BO_ME; ++j) lam[b * (BO_ME + BO_MI) + j] = W[BO_OFF_Y + j];
    for (int i = 0; i < BO_MI; ++i) lam[b * (BO_ME + BO_MI) + BO_ME + i] = W[BO_OFF_Z + i];
    f[b] = S.f; iters[b] = S.it; kkt[b] = S.err0; if (trips) trips[b] = S.trips;
  }
}
extern "C" void hostsim_coop_counts(long long* out, int reset) { for (int k = 0; k < 8; ++k) { out[k] = bo_host_prof[k]; if (reset) bo_host_prof[k] = 0; } }
// component probes for tests/test_coop_logic.py
struct Probe {
  std::vector<double> W, vals, red; int ibuf[4]; bo_cta C;
  Probe(const int* tab, const double* dtab) : W(BO_SCRATCH_DOUBLES), vals(BO_VALS + 1 + BO_NK + 1), red(64) {
    C.W = W.data(); C.vals = vals.data(); C.bp = vals.data() + BO_VALS + 1; C.red = red.data(); C.ibuf = ibuf; C.tab = tab; C.dtab = dtab; C.prof = nullptr; C.wkkt = nullptr; C.wfc = nullptr;
  }
};
static void cp(double* dst, const double* src, int n) { for (int i = 0; i < n; ++i) dst[i] = src[i]; }
extern "C" void hostsim_coop_kkt(const int* tab, const double* dtab, const double* p, const double* x, const double* y, const double* z,
                                 double* f, double* g, double* cE, double* cI, double* JE, double* JI, double* H) {
  Probe P(tab, dtab); double* W = P.W.data();
  cp(W + BO_OFF_P, p, BO_NP); cp(W + BO_OFF_X, x, BO_NX); cp(W + BO_OFF_Y, y, BO_ME); cp(W + BO_OFF_Z, z, BO_MI);
  bo_cta_pre(P.C);
  *f = bo_cta_eval_kkt(P.C);
  cp(g, W + BO_OFF_G, BO_NX); cp(cE, W + BO_OFF_CE, BO_ME); cp(cI, W + BO_OFF_CI, BO_MI);
  cp(JE, W + BO_OFF_JE, BO_NNZ_JE); cp(JI, W + BO_OFF_JI, BO_NNZ_JI); cp(H, W + BO_OFF_H, BO_NNZ_H);
}
extern "C" void hostsim_coop_fc(const int* tab, const double* dtab, const double* p, const double* x, double* f, double* cE, double* cI) {
  Probe P(tab, dtab); double* W = P.W.data();
  cp(W + BO_OFF_P, p, BO_NP); cp(W + BO_OFF_XT, x, BO_NX);
  bo_cta_pre(P.C);
  *f = bo_cta_eval_fc(P.C, W + BO_OFF_XT, W + BO_OFF_CET, W + BO_OFF_CIT);
  cp(cE, W + BO_OFF_CET, BO_ME); cp(cI, W + BO_OFF_CIT, BO_MI);
}
// assemble K(H, JE, JI, sigma, rho) + diag(dw, -dcp), factor, solve K sol = rhs (in place).  Returns the factor's verdict.
// merged != 0: the right-hand side rides through the factor program (forward substitution folded in), then only the tail.
extern "C" int hostsim_coop_linsolve(const int* tab, const double* dtab, const double* H, const double* JE, const double* JI,
                                     const double* sigma, double rho, double dw, double dcp, double* rhs, int merged) {
  Probe P(tab, dtab); double* W = P.W.data();
  cp(W + BO_OFF_H, H, BO_NNZ_H); cp(W + BO_OFF_JE, JE, BO_NNZ_JE); cp(W + BO_OFF_JI, JI, BO_NNZ_JI); cp(W + BO_OFF_SIG, sigma, BO_MI);
  cp(W + BO_OFF_SOL, rhs, BO_NK);
  bo_cta_assemble(P.C, rho, dw, dcp);
  if (merged) bo_cta_load_rhs(P.C, W + BO_OFF_SOL);
  const int bad = bo_cta_factor(P.C);
  if (merged) bo_cta_ldl_solve_tail(P.C, W + BO_OFF_SOL);
  else bo_cta_ldl_solve(P.C, W + BO_OFF_SOL);
  cp(rhs, W + BO_OFF_SOL, BO_NK);
  return bad;
}
"""


class HostSim:
    def __init__(self, generated_source: str, nx: int, np_: int, n_eq: int, n_ineq: int, trace: bool = False, defines: str = "",
                 ldl_table=None, dtable=None):
        self.dtable = None if dtable is None or len(dtable) == 0 else np.ascontiguousarray(dtable, dtype=np.float64)
        self.scratch = np.zeros(int(ldl_table[0] + ldl_table[1]) + 1) if ldl_table is not None and len(ldl_table) else None
        self.ldl_table = None if ldl_table is None or len(ldl_table) == 0 else np.ascontiguousarray(ldl_table, dtype=np.int32)
        self.nx, self.np_, self.nl = nx, np_, n_eq + n_ineq
        key = hashlib.sha1(generated_source.encode()).hexdigest()[:16] + str(trace) + defines
        coop = "#define BO_COOP 1" in generated_source
        wrapper = _WRAPPER_TEAM if "#define BO_TEAM 1" in generated_source else (_WRAPPER_COOP if coop else _WRAPPER)
        key += hashlib.sha1(wrapper.encode()).hexdigest()[:8]
        for name in ("bo_common.cuh", "bo_ipm_reg.cuh", "bo_qp_reg.cuh", "bo_ipm_cta.cuh", "bo_ipm_team.cuh", "bo_team_layout.cuh"):
            key += hashlib.sha1(open(os.path.join(_JIT_INC, name), "rb").read()).hexdigest()[:8]
        d = os.path.join(tempfile.gettempdir(), "b200optas_hostsim")
        os.makedirs(d, exist_ok=True)
        so = os.path.join(d, f"hostsim_{hashlib.sha1(key.encode()).hexdigest()[:16]}.so")
        if not os.path.exists(so):
            gen = os.path.join(d, f"gen_{os.getpid()}.cu")
            wrap = os.path.join(d, f"wrap_{os.getpid()}.cpp")
            open(gen, "w").write(generated_source)
            open(wrap, "w").write(wrapper % {"gen": gen, "trace": ("#define BO_HOST_TRACE 1\n" if trace else "") + defines})
            subprocess.run(["/usr/bin/g++", "-O1", "-shared", "-fPIC", "-std=c++17", "-I", _JIT_INC, "-I", os.path.join(_HERE, "..", "include"), wrap, "-o", so + f".tmp{os.getpid()}"],
                           check=True)
            os.replace(so + f".tmp{os.getpid()}", so)
        self.lib = C.CDLL(so)
        vp = C.c_void_p
        self.lib.hostsim_solve.argtypes = [C.c_longlong] + [vp] * 9 + [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, vp, vp, vp]

    def coop_counts(self, reset=True):
        """cooperative tier only: events since the last reset {kkt tape, f/c tape, assembly, factorisations, solves}"""
        out = np.zeros(8, dtype=np.int64)
        self.lib.hostsim_coop_counts(out.ctypes.data_as(C.c_void_p), int(reset))
        return {"kkt": int(out[0]), "fc": int(out[1]), "assembly": int(out[2]), "factor": int(out[5]), "solve": int(out[6])}

    def solve(self, P, X0, max_iter=100, tol=1e-8, acc_tol=1e-6, mu_init=0.1, max_step=0.5, max_trips=250):
        B = X0.shape[0]
        P = np.ascontiguousarray(P, dtype=float)
        X0 = np.ascontiguousarray(X0, dtype=float)
        X = np.empty((B, self.nx)); lam = np.empty((B, max(self.nl, 1))); f = np.empty(B)
        st = np.empty(B, dtype=np.int32); it = np.empty(B, dtype=np.int32); kkt = np.empty(B); trips = np.empty(B, dtype=np.int32)
        self.lib.hostsim_solve(B, P.ctypes.data, X0.ctypes.data, X.ctypes.data, lam.ctypes.data, f.ctypes.data,
                               st.ctypes.data, it.ctypes.data, kkt.ctypes.data, trips.ctypes.data, max_iter, tol, acc_tol, mu_init, max_step, max_trips,
                               None if self.ldl_table is None else self.ldl_table.ctypes.data,
                               None if self.dtable is None else self.dtable.ctypes.data,
                               None if self.scratch is None else self.scratch.ctypes.data)
        return {"x": X, "lam": lam[:, :self.nl], "f": f, "status": st, "iters": it, "kkt": kkt, "trips": trips}
