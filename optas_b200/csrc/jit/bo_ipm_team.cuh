// bo_ipm_team.cuh -- batched primal-dual interior-point solver for SMALL DENSE problems (nx + n_eq <= 14:
// C1 / C2 inverse kinematics, Booth, differential-IK QPs), ONE INSTANCE PER TEAM OF BO_G THREADS, the whole
// per-instance state in SHARED MEMORY, all iterations inside one launch.
//
// Replaces, like bo_ipm_reg.cuh, what the reference does per call inside CasADiSolver._solve
// (optas/solver.py:386-398 -> casadi nlpsol("ipopt")); same algorithm, same constants, same status codes.
//
// Why this tier exists (round-1 measurements, profiles/r01_solve_details.txt): with one instance per thread the
// iterate, the 10x10 KKT factor and the evaluation results (~3 KB) do not fit 255 registers, live in thread-local
// memory and turn a kernel with 13 MB of algorithmic I/O into one that moves 8.1 GB through DRAM per launch.
//
// Layout of the work
//   * a CTA is BO_G warps = 32 teams; team t = lane t of every warp.  The G threads of a team sit in DIFFERENT
//     warps, so each warp can run its own straight-line code: the expression tapes are cut into BO_G slices by
//     output (bo_team.cpp balances them; shared sub-expressions are recomputed) and warp r evaluates slice r for
//     the 32 instances of the CTA in lock step -- warp-uniform control flow, no divergence, 1/G of the latency.
//     The per-row barrier work of the inequality rows (1/s, log s, |c - s|) rides in the slice that owns the row.
//   * state lives in shared memory as [element][32 lanes]: every access of a warp is 256 contiguous bytes
//     (conflict-free), nothing spills: x, s, y, z, the evaluation (g, c, J, H), the packed LDL' factor, the step
//     and the filter.  ~3.4 KB per instance for C2 => 2 CTAs (64 instances, 8 warps) per SM.
//   * warp 0 is the MASTER of its 32 instances: it owns the scalar state machine in registers (phase, mu, filter
//     sizes ...), assembles and factors the KKT matrix in registers (unrolled, unpivoted LDL' on the rho-augmented
//     system -- see bo_ipm_reg.cuh for why that is valid) and takes every decision; the other warps only ever
//     execute slices, so no decision is computed twice with possibly different rounding.
//   * one trip of the loop = W1 (KKT tape slices, which also assemble the (1,1) block of the KKT matrix and the row
//     sums / maxima of the convergence test) | M1 (residuals, convergence test, mu, LDL', step) | W2a (trial point and
//     its sin / cos: computed once, shared by all slices and by the next iteration's W1) | W2b (f / c slices, logs and
//     reciprocals of the slack rows) | M2 (filter acceptance, SOC, backtracking): five CTA barriers.
//     As in bo_ipm_reg.cuh every data-dependent retry is a state transition that takes effect on the next trip.
//   * code size matters as much as instruction count: every warp runs its own straight-line code, so the instruction
//     working set of an SM is the whole kernel.  The first version (250 KB of SASS: sin / cos and log expanded in every
//     slice) spent 17 of every 43 cycles per issued instruction waiting for instruction fetch
//     (profiles/r02_team_v0_ncu.txt); transcendental functions now exist once (non-inlined helpers, shared sin / cos),
//     loops with a division or a logarithm in the body are not unrolled, constants come from the constant bank.
//   * persistent CTAs; a team that finishes fetches the next instance (global counter) -- with TMA staging:
//     the CTA pulls TILES of 32 consecutive instances (p and x0 rows are contiguous byte ranges) into shared memory
//     with cp.async.bulk on an mbarrier, one tile ahead, and hands them out one by one.
//
// Generated prelude (bo_team.cpp) defines BO_NX, BO_NP, BO_ME, BO_MI, BO_NNZ_JE, BO_NNZ_JI, BO_NNZ_H, BO_G, BO_TPB,
// the tape slices bo_team_kkt(role, sm) / bo_team_fc(role, sm, at, rows) and the strided sparse helpers *_t.
#pragma once
#include "bo_common.cuh"

#include "bo_team_layout.cuh"

// Master-only scalar state of one instance (registers of warp 0).
struct bo_tm {
  double f, mu, tau, dw_last, err0, theta_max, theta_min, phi0, theta0, dw, dc, rho, a, a_trial, dphi, th_soc;
  double a_ftype;  // step lengths above this make the switching condition of the filter hold for the current direction
  double lgs;      // sum log(s) at the current iterate (from the slices that evaluated the accepted trial point)
  double thr, thr0;  // feasibility restoration: l1 infeasibility of (c_E, min(c_I, 0)) at x, and at entry
  int nf, it, n_acceptable, phase, trips, attempt, heavy, n_singular, ls, soc;
  int resto, n_resto;  // iterations spent in the current restoration phase (0: regular mode); phases entered so far
  bool recalc_y, ls_mode, jac_degenerate, first_singular;
  long long b;
};

// ---- feasibility restoration (stands in for IPOPT's restoration phase, Waechter & Biegler section 3.3) ----
// Entered when the line search finds no acceptable step at an infeasible point (typically: the fraction-to-the-boundary
// rule leaves a step length of 1e-6 because a slack sits on its bound while its constraint is violated).  The iterate
// then moves by Levenberg-Marquardt steps on the infeasibility alone,
//     min_dx  1/2 || c_E + J_E dx ||^2 + 1/2 || min(c_I + J_I dx, 0) ||^2 + zeta/2 ||dx||^2
// through the same KKT factorisation (H := zeta I, Sigma := indicator of the violated rows, z := 0, mu := 0, rho := 1 and
// the constraint block made inert by dc -> 1), with an Armijo test on theta_r = || (c_E, min(c_I, 0)) ||_2; a step
// that fails it twice is recomputed with 100 x the damping zeta.  It ends when theta_r has dropped to BO_RESTO_KAPPA of its
// entry value: slacks and multipliers are re-initialised at the new point (as for a fresh instance, mu kept) and the
// regular iteration resumes; or with BO_ST_LINE_SEARCH when theta_r cannot be reduced (a stationary point of the
// infeasibility: what IPOPT reports as "converged to a point of local infeasibility").
#ifndef BO_RESTO_KAPPA
#define BO_RESTO_KAPPA 0.1
#endif
#ifndef BO_RESTO_ZETA
#define BO_RESTO_ZETA 1e-4
#endif
#ifndef BO_RESTO_MAX_IT
#define BO_RESTO_MAX_IT 40
#endif
#ifndef BO_RESTO_MAX_PHASES
#define BO_RESTO_MAX_PHASES 3
#endif
#define BO_RESTO_DC (1.0 - 1e-8) /* with rho = 1: constraint block -dc / (1 - rho dc) = -1e8, i.e. inert */
#ifndef BO_ALPHA_MIN
#define BO_ALPHA_MIN 5e-7 /* IPOPT's alpha_min = gamma_alpha gamma_theta for a non-descent direction */
#endif

#ifdef BO_HOST_SIM
#define BO_CTZLL(m) __builtin_ctzll(m)
static inline double bo_rcp(double d) { return 1.0 / d; }
static double bo_log_ni(double v) { return std::log(v); }
static double bo_exp_ni(double v) { return std::exp(v); }
#else
#define BO_CTZLL(m) (__ffsll((long long)(m)) - 1)
// Reciprocal without the IEEE slow path: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps, within 2 ulp for
// normal arguments.  ~6 instructions instead of ~40, and the factorisation alone needs one per pivot.
__device__ __forceinline__ double bo_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  if (r != 0.0 && fabs(r) < BO_INF) {  // seed 0 / inf (huge, infinite, zero or denormal argument) and NaN are final
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
  }
  return r;
}
// one copy of the logarithm / exponential (~100 instructions each when inlined) for all their call sites
__device__ __noinline__ double bo_log_ni(double v) { return log(v); }
__device__ __noinline__ double bo_exp_ni(double v) { return exp(v); }
#endif

// ---- W2a: trial point and the shared sin / cos of its components (one copy of this code serves every role) ----
BO_NOINLINE void bo_team_pre(double* BO_RESTRICT sm, const double at, const int role) {
  BO_NOUNROLL
  for (int k = role; k < BO_NX; k += BO_G) {
    const double xt = SM(BO_OFF_X, k) + at * SM(BO_OFF_DX, k);
    SM(BO_OFF_XT, k) = xt;
    if ((BO_TRIG_MASK >> k) & 1u) {
      double sn, cs;
      bo_sincos(xt, &sn, &cs);
      SM(BO_OFF_SH, 2 * k) = sn;
      SM(BO_OFF_SH, 2 * k + 1) = cs;
    }
  }
}

// ---- W2b tail: per-row barrier work of the rows a slice owns (its f / c outputs are already in shared memory) ----
// trial slack st = s + a ds, log st, |c - st|; and, in case the step is accepted, the reset slack max(st, c), its
// reciprocal and its log -- so that the iteration that follows needs no division and no logarithm at all.
BO_NOINLINE void bo_team_rows(double* BO_RESTRICT sm, double at, bool rows, int role, unsigned long long mask_e,
                              unsigned long long mask_i) {
  double lg = 0.0, th = 0.0, lgn = 0.0;
  while (mask_e) {
    const int j = BO_CTZLL(mask_e);
    mask_e &= mask_e - 1;
    th += fabs(SM(BO_OFF_CET, j));
  }
  if (rows) {
    while (mask_i) {
      const int i = BO_CTZLL(mask_i);
      mask_i &= mask_i - 1;
      const double st = SM(BO_OFF_S, i) + at * SM(BO_OFF_DS, i);
      const double cit = SM(BO_OFF_CIT, i);
      const double l = bo_log_ni(st);
      lg += l;
      th += fabs(cit - st);
      const double sn = fmax(st, cit);  // slack reset: lowers theta, never raises the barrier objective
      const double rsn = bo_rcp(sn);
      SM(BO_OFF_SN, i) = sn;
      SM(BO_OFF_RSN, i) = rsn;
      // log(sn) = log(st) - log(1 - d), d = (sn - st) / sn >= 0.  For linear rows c and st differ by rounding only: d is
      // 0 or ~1e-16 and the first-order term is exact to double precision; a second logarithm only for a real reset.
      const double d = (sn - st) * rsn;
      lgn += (d < 1e-9) ? l + d : bo_log_ni(sn);
    }
  }
  SM(BO_OFF_PART, 5 * role) = lg;
  SM(BO_OFF_PART, 5 * role + 1) = th;
  SM(BO_OFF_PART, 5 * role + 2) = lgn;
}

// ---- W1 tail: the part of the convergence test that is a sum / max over the constraint rows a slice owns ----
// also leaves the linearisation residuals RE = cE, RI = cI - s for the step computation
BO_NOINLINE void bo_team_rows_kkt(double* BO_RESTRICT sm, int role, unsigned long long mask_e, unsigned long long mask_i) {
  double e_prim = 0.0, theta = 0.0, sz_min = BO_INF, sz_max = 0.0, sum_z = 0.0;
  while (mask_e) {
    const int j = BO_CTZLL(mask_e);
    mask_e &= mask_e - 1;
    const double c = SM(BO_OFF_CE, j);
    SM(BO_OFF_RE, j) = c;
    e_prim = fmax(e_prim, fabs(c));
    theta += fabs(c);
  }
  while (mask_i) {
    const int i = BO_CTZLL(mask_i);
    mask_i &= mask_i - 1;
    const double s = SM(BO_OFF_S, i), z = SM(BO_OFF_Z, i);
    const double r = SM(BO_OFF_CI, i) - s;
    SM(BO_OFF_RI, i) = r;
    e_prim = fmax(e_prim, fabs(r));
    theta += fabs(r);
    sz_min = fmin(sz_min, s * z);
    sz_max = fmax(sz_max, s * z);
    sum_z += fabs(z);
  }
  SM(BO_OFF_PART, 5 * role) = e_prim;
  SM(BO_OFF_PART, 5 * role + 1) = theta;
  SM(BO_OFF_PART, 5 * role + 2) = sz_min;
  SM(BO_OFF_PART, 5 * role + 3) = sz_max;
  SM(BO_OFF_PART, 5 * role + 4) = sum_z;
}

// ---- unpivoted LDL' of the packed matrix in registers (see bo_ipm_reg.cuh: valid on the rho-augmented system) ----
// On exit A holds 1/D on the diagonal and the unit-lower L below it.
BO_DEVICE int bo_tm_ldl(double* BO_RESTRICT A) {
  int bad = 0;  // 0 ok, 1 = non-positive pivot in the x block, 2 = non-negative pivot in the y block
  BO_UNROLL
  for (int j = 0; j < BO_NK; ++j) {
    double d = A[BO_KIDX(j, j)];
    const double scale = fmax(1.0, fabs(d));
    BO_UNROLL
    for (int k = 0; k < j; ++k) {
      const double c = A[BO_KIDX(j, k)];  // still the unscaled C(j,k) = L(j,k) D(k)
      d -= c * c * A[BO_KIDX(k, k)];      // A(k,k) already 1/D(k)
    }
    if (j < BO_NX) {
      if (!(d > 1e-13 * scale) && bad == 0) bad = 1;
    } else {
      if (!(d < -1e-13) && bad == 0) bad = 2;
    }
    const double dinv = bo_rcp(d);
    BO_UNROLL
    for (int i = j + 1; i < BO_NK; ++i) {
      double v = A[BO_KIDX(i, j)];
      BO_UNROLL
      for (int k = 0; k < j; ++k) v -= A[BO_KIDX(i, k)] * A[BO_KIDX(j, k)] * A[BO_KIDX(k, k)];
      A[BO_KIDX(i, j)] = v;  // unscaled C(i,j) = L(i,j) D(j): L(i,k) L(j,k) D(k) = C(i,k) C(j,k) / D(k)
    }
    A[BO_KIDX(j, j)] = dinv;
  }
  // C -> L
  BO_UNROLL
  for (int j = 0; j < BO_NK; ++j) {
    BO_UNROLL
    for (int i = j + 1; i < BO_NK; ++i) A[BO_KIDX(i, j)] *= A[BO_KIDX(j, j)];
  }
  return bad;
}

// Solve K sol = sol in place on the SOL vector in shared memory, with the factor stored there too (1/D on the
// diagonal).  Fully unrolled: the vector is loaded once, lives in registers, and is stored once.
BO_NOINLINE void bo_tm_ldl_solve(double* BO_RESTRICT sm) {
  double b[BO_NK];
  BO_UNROLL
  for (int i = 0; i < BO_NK; ++i) b[i] = SM(BO_OFF_SOL, i);
  BO_UNROLL
  for (int i = 1; i < BO_NK; ++i) {
    BO_UNROLL
    for (int k = 0; k < i; ++k) b[i] -= SM(BO_OFF_LD, BO_KIDX(i, k)) * b[k];
  }
  BO_UNROLL
  for (int i = 0; i < BO_NK; ++i) b[i] *= SM(BO_OFF_LD, BO_KIDX(i, i));
  BO_UNROLL
  for (int i = BO_NK - 2; i >= 0; --i) {
    BO_UNROLL
    for (int k = i + 1; k < BO_NK; ++k) b[i] -= SM(BO_OFF_LD, BO_KIDX(k, i)) * b[k];
  }
  BO_UNROLL
  for (int i = 0; i < BO_NK; ++i) SM(BO_OFF_SOL, i) = b[i];
}

// Step for the constraint residuals (RE, RI) with the current factorisation: writes DX, DS (and, when y_step is
// set, YST) and returns the fraction-to-the-boundary primal step length; SOL[0] carries sum ds / s back.
BO_NOINLINE double bo_tm_step(double* BO_RESTRICT sm, const double mu, const double rho, const double dc, const double tau,
                              const bool y_step) {
  {
    double sol[BO_NX];
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) sol[i] = -SM(BO_OFF_RD, i);
    // t = -(z - mu / s + sigma * rI), parked in DS (overwritten by the real ds below)
    BO_NOUNROLL
    for (int i = 0; i < BO_MI; ++i)
      SM(BO_OFF_DS, i) = -(SM(BO_OFF_Z, i) - mu * SM(BO_OFF_RS, i) + SM(BO_OFF_SIG, i) * SM(BO_OFF_RI, i));
    bo_JIt_acc_t(SMP(BO_OFF_JI), SMP(BO_OFF_DS), BO_LS, 1.0, sol, 1);
    bo_JEt_acc_t(SMP(BO_OFF_JE), SMP(BO_OFF_RE), BO_LS, -rho, sol, 1);  // first block row += rho JE' (second block rhs)
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_SOL, i) = sol[i];
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_SOL, BO_NX + j) = -SM(BO_OFF_RE, j);
  }
  bo_tm_ldl_solve(sm);
  BO_UNROLL
  for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_DX, i) = SM(BO_OFF_SOL, i);
  if (y_step) {
    const double undo = bo_rcp(1.0 - rho * dc);
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_YST, j) = -SM(BO_OFF_SOL, BO_NX + j) * undo;
  }
  bo_JI_mul_t(SMP(BO_OFF_JI), SMP(BO_OFF_SOL), BO_LS, SMP(BO_OFF_DS), BO_LS);
  double worst = 0.0, dsr = 0.0;  // max over rows of -ds / s; sum ds / s (directional derivative of the barrier term)
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) {
    const double ds = SM(BO_OFF_DS, i) + SM(BO_OFF_RI, i);
    SM(BO_OFF_DS, i) = ds;
    const double q = ds * SM(BO_OFF_RS, i);
    worst = fmax(worst, -q);
    dsr += q;
  }
  SM(BO_OFF_SOL, 0) = dsr;  // the solution vector is consumed: its first slot carries sum ds / s back to the caller
  return worst > tau ? tau / worst : 1.0;  // min(1, min_i -tau s_i / ds_i)
}

// Least-squares multiplier estimate (bo_ipm_reg.cuh): with [I JE'; JE -dc] factored, solve for [r; y] = K \ [g - JI'z; 0],
// optionally followed by one step of iterative refinement towards the unregularised solution.
BO_NOINLINE void bo_tm_ls_multipliers(double* BO_RESTRICT sm, const double dw, const bool refine) {
  {
    double sol[BO_NX];
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) sol[i] = SM(BO_OFF_G, i);
    bo_JIt_acc_t(SMP(BO_OFF_JI), SMP(BO_OFF_Z), BO_LS, -1.0, sol, 1);
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_SOL, i) = sol[i];
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_SOL, BO_NX + j) = 0.0;
  }
  bo_tm_ldl_solve(sm);
  bool fin = true;
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) fin = fin && bo_isfinite(SM(BO_OFF_SOL, BO_NX + j));
  if (!fin) return;
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_Y, j) = SM(BO_OFF_SOL, BO_NX + j);
  if (!refine) return;
  {
    double sol[BO_NX];
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) {
      const double r = SM(BO_OFF_SOL, i);
      SM(BO_OFF_DX0, i) = r;
      sol[i] = SM(BO_OFF_G, i) - dw * r;
    }
    bo_JIt_acc_t(SMP(BO_OFF_JI), SMP(BO_OFF_Z), BO_LS, -1.0, sol, 1);
    bo_JEt_acc_t(SMP(BO_OFF_JE), SMP(BO_OFF_Y), BO_LS, -1.0, sol, 1);
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_SOL, i) = sol[i];
    // second block: -(JE r), formed in shared memory (SOL tail), read back negated
    bo_JE_mul_t(SMP(BO_OFF_JE), SMP(BO_OFF_DX0), BO_LS, SMP(BO_OFF_SOL) + BO_NX * BO_LS, BO_LS);
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_SOL, BO_NX + j) = -SM(BO_OFF_SOL, BO_NX + j);
  }
  bo_tm_ldl_solve(sm);
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) {
    const double d = SM(BO_OFF_SOL, BO_NX + j);
    if (bo_isfinite(d)) SM(BO_OFF_Y, j) += d;
  }
}

// Restoration step data at x: weights / residuals of the violated inequality rows in SIG / RI, no barrier, no multipliers.
// Returns theta_r.  (S, Z, RS, SIG are re-initialised when the restoration phase ends.)
BO_NOINLINE double bo_tm_resto_prepare(double* BO_RESTRICT sm) {
  double thr = 0.0;
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) {
    const double c = SM(BO_OFF_CE, j);
    SM(BO_OFF_RE, j) = c;
    thr += c * c;
  }
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) {
    const double v = fmin(SM(BO_OFF_CI, i), 0.0);
    SM(BO_OFF_SIG, i) = v < 0.0 ? 1.0 : 0.0;
    SM(BO_OFF_RI, i) = v;
    SM(BO_OFF_Z, i) = 0.0;
    SM(BO_OFF_RS, i) = 0.0;
    thr += v * v;
  }
  BO_UNROLL
  for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_RD, i) = 0.0;
  return sqrt(thr);
}

// Slacks pushed into the interior, z on the central path, y = 0, x := XT: start of an instance and end of a restoration
// phase.  Needs c_I at XT in CIT.
BO_NOINLINE double bo_tm_init_point(double* BO_RESTRICT sm, const double mu) {
  double lgs = 0.0;
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) {
    const double ci = SM(BO_OFF_CIT, i);
    const double s = fmax(ci, 1e-2 * fmax(1.0, fabs(ci)));
    const double rs = bo_rcp(s);
    SM(BO_OFF_S, i) = s;
    SM(BO_OFF_RS, i) = rs;
    SM(BO_OFF_Z, i) = mu * rs;
    SM(BO_OFF_SIG, i) = mu * rs * rs;
    lgs += bo_log_ni(s);
  }
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_Y, j) = 0.0;
  BO_UNROLL
  for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_X, i) = SM(BO_OFF_XT, i);
  return lgs;
}

// A fresh instance: the master has put p and the seed into P / X.
BO_DEVICE void bo_tm_begin(bo_tm& M, double* BO_RESTRICT sm, const bo_solver_params& prm, long long b) {
  M.b = b;
  M.mu = prm.mu_init;
  M.dw_last = 0.0;
  M.err0 = BO_INF;
  M.theta_max = BO_INF;
  M.theta_min = 0.0;
  M.nf = 0;
  M.it = 0;
  M.n_acceptable = 0;
  M.recalc_y = false;
  M.ls_mode = false;
  M.n_singular = 0;
  M.jac_degenerate = false;
  M.first_singular = false;
  M.trips = 0;
  M.f = 0.0;
  M.a_trial = 0.0;
  M.a_ftype = BO_INF;
  M.lgs = 0.0;
  M.resto = 0;
  M.n_resto = 0;
  M.thr = 0.0;
  M.thr0 = 0.0;
  M.phase = BO_PH_INIT;
  BO_UNROLL
  for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_DX, i) = 0.0;  // the INIT evaluation reads x + 0 * dx
  SM(BO_OFF_AT, 0) = 0.0;
}

// ---- M1: EVAL part (results of the KKT slices are in shared memory) and FACTOR part ----
// Returns -1 to continue or the final status.
BO_DEVICE int bo_tm_m1(bo_tm& M, double* BO_RESTRICT sm, const bo_solver_params& prm) {
  const double kappa_eps = 10.0, kappa_mu = 0.2, tau_min = 0.99, s_max = 100.0, s_phi = 2.3, s_theta = 1.1;
  const double mu_min = prm.tol * 0.1;
  const bool over = ++M.trips > prm.max_trips;
  if (over && M.phase != BO_PH_EVAL) return BO_ST_MAX_ITER;

  if (M.phase == BO_PH_EVAL) {
    M.f = SM(BO_OFF_F0, 0);
    if (M.resto > 0) {
      // restoration: next Levenberg-Marquardt step on the infeasibility from the fresh Jacobians
      if (!bo_isfinite(M.f)) return BO_ST_NUMERICAL;
      if (over || M.it >= prm.max_iter) return BO_ST_MAX_ITER;
      if (M.resto > BO_RESTO_MAX_IT) return BO_ST_LINE_SEARCH;
      M.thr = bo_tm_resto_prepare(sm);
      M.dw = BO_RESTO_ZETA;
      M.dc = BO_RESTO_DC;
      M.first_singular = false;
      M.attempt = 0;
      M.heavy = 0;
      M.ls_mode = false;
      M.phase = BO_PH_FACTOR;
    } else if (M.recalc_y && BO_ME > 0 && !over) {
      // least-squares multiplier estimate after a regularised step (bo_ipm_reg.cuh): factor [I JE'; JE -dc]
      M.recalc_y = false;
      M.ls_mode = true;
      M.dw = 1.0;
      M.dc = 1e-10;
      M.phase = BO_PH_FACTOR;
    } else {
      double e_dual = 0.0, e_prim = 0.0, e_comp0 = 0.0, sum_mult = 0.0, sum_z = 0.0, theta = 0.0, sz_min = BO_INF;
      {
        double rd[BO_NX];
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) rd[i] = SM(BO_OFF_G, i);
        bo_JEt_acc_t(SMP(BO_OFF_JE), SMP(BO_OFF_Y), BO_LS, -1.0, rd, 1);
        bo_JIt_acc_t(SMP(BO_OFF_JI), SMP(BO_OFF_Z), BO_LS, -1.0, rd, 1);
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) {
          SM(BO_OFF_RD, i) = rd[i];
          e_dual = fmax(e_dual, fabs(rd[i]));
        }
      }
      BO_UNROLL
      for (int r = 0; r < BO_G; ++r) {  // partial results of the slices' row work
        e_prim = fmax(e_prim, SM(BO_OFF_PART, 5 * r));
        theta += SM(BO_OFF_PART, 5 * r + 1);
        sz_min = fmin(sz_min, SM(BO_OFF_PART, 5 * r + 2));
        e_comp0 = fmax(e_comp0, SM(BO_OFF_PART, 5 * r + 3));
        sum_z += SM(BO_OFF_PART, 5 * r + 4);
      }
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) sum_mult += fabs(SM(BO_OFF_Y, j));
      sum_mult += sum_z;
      // scaling of the dual / complementarity errors (Waechter & Biegler eq. 6), as reciprocals: 1 / s_d, 1 / s_c
      const double rs_d = (BO_ME + BO_MI) > 0 ? bo_rcp(fmax(1.0, sum_mult * (1.0 / (s_max * (double)BO_DIM(BO_ME + BO_MI))))) : 1.0;
      const double rs_c = BO_MI > 0 ? bo_rcp(fmax(1.0, sum_z * (1.0 / (s_max * (double)BO_DIM(BO_MI))))) : 1.0;
      M.err0 = fmax(fmax(e_dual * rs_d, e_prim), e_comp0 * rs_c);
#ifdef BO_HOST_TRACE
      printf("it %3d f %.6e err0 %.3e (dual %.3e prim %.3e comp %.3e) mu %.2e nf %d dw_last %.2e\n", M.it, M.f, M.err0,
             e_dual * rs_d, e_prim, e_comp0 * rs_c, M.mu, M.nf, M.dw_last);
#endif
      if (!bo_isfinite(M.err0) || !bo_isfinite(M.f)) return BO_ST_NUMERICAL;
      if (M.err0 <= prm.tol) return BO_ST_CONVERGED;
      M.n_acceptable = (M.err0 <= prm.acceptable_tol) ? M.n_acceptable + 1 : 0;
      if (M.n_acceptable >= 15) return BO_ST_ACCEPTABLE;
      if (M.it >= prm.max_iter || over) return BO_ST_MAX_ITER;

      // barrier parameter update (monotone Fiacco-McCormick, Waechter & Biegler eq. 7); resets the filter
      if (BO_MI > 0) {
        BO_NOUNROLL
        for (int rep = 0; rep < 8; ++rep) {
          const double e_comp = fmax(e_comp0 - M.mu, M.mu - sz_min);  // max_i |s_i z_i - mu|
          const double err_mu = fmax(fmax(e_dual * rs_d, e_prim), e_comp * rs_c);
          if (err_mu <= kappa_eps * M.mu && M.mu > mu_min) {
            M.mu = fmax(mu_min, fmin(kappa_mu * M.mu, M.mu * sqrt(M.mu)));
            M.nf = 0;
          } else {
            break;
          }
        }
      }
      M.tau = fmax(tau_min, 1.0 - M.mu);
      M.theta0 = theta;
      M.phi0 = M.f - M.mu * M.lgs;
      if (M.it == 0) {
        M.theta_max = 1e4 * fmax(1.0, M.theta0);
        M.theta_min = 1e-4 * fmax(1.0, M.theta0);
      }
      M.dw = 0.0;
      M.dc = M.jac_degenerate ? BO_DC_SCALE * sqrt(sqrt(M.mu)) : 0.0;  // IPOPT's degeneracy heuristic
      M.first_singular = false;
      M.attempt = 0;
      M.heavy = 0;
      M.ls_mode = false;
      M.phase = BO_PH_FACTOR;
    }
  }

  if (M.phase == BO_PH_FACTOR) {
    const double rho = M.ls_mode ? 0.0 : (M.resto > 0 ? 1.0 : BO_STATIC_RHO);
    M.rho = rho;
    int bad;
    {
      // K = [ KX + dw I , JE' ; JE , -dc' I ]: the (1,1) block comes assembled from the KKT slices
      double K[BO_KSZ];
      BO_UNROLL
      for (int i = 0; i < BO_KSZ; ++i) K[i] = 0.0;
      if (M.resto > 0) bo_kkt_gn_t(SMP(BO_OFF_JE), SMP(BO_OFF_JI), SMP(BO_OFF_SIG), rho, SMP(BO_OFF_KX));  // JI' W JI + rho JE'JE
      if (!M.ls_mode) {
        BO_UNROLL
        for (int i = 0; i < (BO_NX * (BO_NX + 1)) / 2; ++i) K[i] = SM(BO_OFF_KX, i);
      }
      bo_kkt_je_t(SMP(BO_OFF_JE), K);
      const double dcp = M.dc / (1.0 - rho * M.dc);
      BO_UNROLL
      for (int i = 0; i < BO_NX; ++i) K[BO_KIDX(i, i)] += M.dw;
      BO_UNROLL
      for (int i = BO_NX; i < BO_NK; ++i) K[BO_KIDX(i, i)] -= dcp;
      bad = bo_tm_ldl(K);
      BO_UNROLL
      for (int i = 0; i < BO_KSZ; ++i) SM(BO_OFF_LD, i) = K[i];
    }
    const int inertia = bad == 0 ? 0 : (bad == 1 ? 1 : -1);
    if (M.ls_mode) {
      if (inertia == 0) bo_tm_ls_multipliers(sm, M.dw, M.err0 < BO_REFINE_BELOW);
      M.ls_mode = false;
      M.phase = BO_PH_EVAL;  // re-evaluate the Hessian with the new multipliers on the next trip
      return -1;
    }
#ifdef BO_HOST_TRACE
    if (inertia != 0) printf("     inertia %d at dw %.3e dc %.3e\n", inertia, M.dw, M.dc);
#endif
    if (inertia != 0) {
      // inertia correction (IPOPT Algorithm IC); retried on the next trip
      if (inertia < 0 && BO_ME > 0 && M.dc == 0.0) {
        M.dc = BO_DC_SCALE * sqrt(sqrt(M.mu));
        if (M.attempt == 0 || BO_SINGULAR_ANY_ATTEMPT) M.first_singular = true;
      } else if (M.dw == 0.0) {
        M.dw = (M.dw_last == 0.0) ? 1e-4 : fmax(1e-20, M.dw_last / 3.0);
      } else {
        M.dw *= (M.dw_last == 0.0) ? 100.0 : 8.0;
      }
      if (++M.attempt > BO_IC_MAX || M.dw > 1e40) return M.err0 <= prm.acceptable_tol ? BO_ST_ACCEPTABLE : BO_ST_NUMERICAL;
      return -1;
    }
    if (M.dw > 0.0 && M.heavy == 0 && M.resto == 0) M.dw_last = M.dw;
    if (M.heavy == 0 && !M.jac_degenerate) {
      M.n_singular = M.first_singular ? M.n_singular + 1 : 0;
      if (M.n_singular >= 3) M.jac_degenerate = true;
    }
    if (M.heavy > 0 && M.resto == 0) {
      // RE / RI were left by the KKT slices; the second-order corrections of an earlier direction overwrote them
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_RE, j) = SM(BO_OFF_CE, j);
      BO_NOUNROLL
      for (int i = 0; i < BO_MI; ++i) SM(BO_OFF_RI, i) = SM(BO_OFF_CI, i) - SM(BO_OFF_S, i);
    }
    const double a_p = bo_tm_step(sm, M.resto > 0 ? 0.0 : M.mu, M.rho, M.dc, M.tau, true);
    const double dsr = SM(BO_OFF_SOL, 0);
    double dphi = 0.0, dxn = 0.0;
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) {
      const double dx = SM(BO_OFF_DX, i);
      dphi += SM(BO_OFF_G, i) * dx;
      dxn = fmax(dxn, fabs(dx));
    }
    M.dphi = dphi - M.mu * dsr;
    // Switching condition (Waechter & Biegler eq. 19): a (-dphi)^s_phi > theta0^s_theta, kept as a threshold on a so that
    // the trial points of this direction need no logarithm
    if (M.dphi < 0.0 && M.theta0 <= M.theta_min)
      M.a_ftype = M.theta0 <= 0.0 ? 0.0 : bo_exp_ni(s_theta * bo_log_ni(M.theta0) - s_phi * bo_log_ni(-M.dphi));
    else
      M.a_ftype = BO_INF;
    M.a = a_p;
    if (prm.max_step > 0.0 && M.a * dxn > prm.max_step) M.a = prm.max_step / dxn;
    M.a_trial = M.a;
    M.ls = 0;
    M.soc = 0;
    M.phase = BO_PH_TRIAL;
  }
  return -1;
}

// ---- M2: the trial point has been evaluated by the f / c slices ----
BO_DEVICE int bo_tm_m2(bo_tm& M, double* BO_RESTRICT sm, const bo_solver_params& prm) {
  const double kappa_sigma = 1e10, gamma_theta = 1e-5, gamma_phi = 1e-5, eta_phi = 1e-8;
  const double kappa_soc = 0.99;
  if (M.phase == BO_PH_INIT) {
    // start of an instance: slacks from c_I(x0) pushed into the interior, z on the central path, y = 0
    M.f = SM(BO_OFF_FT, 0);
    M.lgs = bo_tm_init_point(sm, M.mu);
    M.phase = BO_PH_EVAL;
    return -1;
  }
  if (M.phase != BO_PH_TRIAL) return -1;
  const double at = M.a_trial;
  if (M.resto > 0) {
    // restoration trial point: Armijo on theta_r
    double thr_t = 0.0;
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) thr_t += SM(BO_OFF_CET, j) * SM(BO_OFF_CET, j);
    BO_NOUNROLL
    for (int i = 0; i < BO_MI; ++i) {
      const double v = fmin(SM(BO_OFF_CIT, i), 0.0);
      thr_t += v * v;
    }
    thr_t = sqrt(thr_t);
#ifdef BO_HOST_TRACE
    printf("     resto %d ls %d a %.3e theta_r %.3e -> %.3e (entry %.3e)\n", M.resto, M.ls, at, M.thr, thr_t, M.thr0);
#endif
    if (bo_isfinite(thr_t) && thr_t <= (1.0 - 1e-4 * at) * M.thr) {
      M.it += 1;
      if (thr_t <= fmax(BO_RESTO_KAPPA * M.thr0, 1e-10)) {
        // feasible enough: back to the regular iteration from here, multipliers and filter start afresh
        M.f = SM(BO_OFF_FT, 0);
        M.lgs = bo_tm_init_point(sm, M.mu);
        M.resto = 0;
        M.nf = 0;
        M.dw_last = 0.0;
        M.recalc_y = BO_ME > 0;  // least-squares equality multipliers at the new point
      } else {
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_X, i) = SM(BO_OFF_XT, i);
        M.resto += 1;
      }
      M.phase = BO_PH_EVAL;
      return -1;
    }
    if (M.ls >= 2 && M.attempt < 4) {  // Levenberg-Marquardt: more damping, new direction
      M.attempt += 1;
      M.dw *= 100.0;
      M.phase = BO_PH_FACTOR;
      return -1;
    }
    M.a *= 0.5;
    M.a_trial = M.a;
    M.ls += 1;
    if (M.ls >= 24 || M.a < 1e-10) return BO_ST_LINE_SEARCH;  // stationary point of the infeasibility
    return -1;
  }
  const double ft = SM(BO_OFF_FT, 0);
  double lg = 0.0, thetat = 0.0, lgn = 0.0;
  BO_UNROLL
  for (int r = 0; r < BO_G; ++r) {
    lg += SM(BO_OFF_PART, 5 * r);
    thetat += SM(BO_OFF_PART, 5 * r + 1);
    lgn += SM(BO_OFF_PART, 5 * r + 2);
  }
  const double phit = ft - M.mu * lg;
  const bool finite = bo_isfinite(phit) && bo_isfinite(thetat);
  const bool ftype = M.a > M.a_ftype;
  const double slack = 10.0 * 2.2e-16 * fabs(M.phi0);
  bool ok = false, armijo = false;
  if (finite && thetat <= M.theta_max) {
    bool in_filter = true;
    for (int j = 0; j < M.nf; ++j) {
      const double fth = SM(BO_OFF_FTH, j), fph = SM(BO_OFF_FPH, j);
      if (!(thetat <= (1.0 - gamma_theta) * fth || phit <= fph - gamma_phi * fth)) in_filter = false;
    }
    if (in_filter) {
      if (ftype) {
        armijo = phit - M.phi0 - slack <= eta_phi * M.a * M.dphi;
        ok = armijo;
      } else {
        ok = thetat <= (1.0 - gamma_theta) * M.theta0 || phit - slack <= M.phi0 - gamma_phi * M.theta0;
      }
    }
  }
#ifdef BO_HOST_TRACE
  printf("     heavy %d ls %d soc %d a %.3e ok %d ftype %d theta %.3e->%.3e phi %.8e->%.8e dphi %.3e dw %.2e\n", M.heavy, M.ls,
         M.soc, M.a_trial, (int)ok, (int)ftype, M.theta0, thetat, M.phi0, phit, M.dphi, M.dw);
#endif
  if (ok) {
    if (!(ftype && armijo)) {  // augment the filter (eq. 22)
      int slot = M.nf;
      if (M.nf < BO_NFILTER) {
        ++M.nf;
      } else {  // full: overwrite the entry with the largest theta
        slot = 0;
        for (int j = 1; j < BO_NFILTER; ++j)
          if (SM(BO_OFF_FTH, j) > SM(BO_OFF_FTH, slot)) slot = j;
      }
      SM(BO_OFF_FTH, slot) = (1.0 - gamma_theta) * M.theta0;
      SM(BO_OFF_FPH, slot) = M.phi0 - gamma_phi * M.theta0;
    }
    // dual step with its own fraction-to-the-boundary rule: a_d = min(1, tau * min_i z_i / (-dz_i)); the minimum
    // of the ratios is tracked by cross-multiplication (one division in total)
    double num = 1.0, den = 0.0;  // best ratio num / den (den > 0), none yet
    BO_NOUNROLL
    for (int i = 0; i < BO_MI; ++i) {
      const double z = SM(BO_OFF_Z, i);
      const double dz = -z + M.mu * SM(BO_OFF_RS, i) - SM(BO_OFF_SIG, i) * SM(BO_OFF_DS, i);
      SM(BO_OFF_RI, i) = dz;  // RI is dead once the step is accepted
      if (dz < 0.0 && (den == 0.0 || z * den < num * (-dz))) {
        num = z;
        den = -dz;
      }
    }
    const double a_d = (den > 0.0 && M.tau * num < den) ? M.tau * num / den : 1.0;
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_X, i) = SM(BO_OFF_XT, i);
    BO_NOUNROLL
    for (int i = 0; i < BO_MI; ++i) {
      const double s = SM(BO_OFF_SN, i), rs = SM(BO_OFF_RSN, i);  // reset slack and its reciprocal, from the slices
      SM(BO_OFF_S, i) = s;
      SM(BO_OFF_RS, i) = rs;
      double z = SM(BO_OFF_Z, i) + a_d * SM(BO_OFF_RI, i);
      // keep z within a factor kappa_sigma of the central-path value mu/s (IPOPT eq. 16); the test is division-free
      const double sz = s * z;
      if (sz > kappa_sigma * M.mu) z = kappa_sigma * M.mu * rs;
      else if (sz * kappa_sigma < M.mu) z = M.mu * rs / kappa_sigma;
      SM(BO_OFF_Z, i) = z;
      SM(BO_OFF_SIG, i) = z * rs;
    }
    M.lgs = lgn;
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_Y, j) += M.a * SM(BO_OFF_YST, j);
#ifdef BO_RECALC_DC_ONLY
    M.recalc_y = M.dc > 0.0;
#else
    M.recalc_y = M.dw > 0.0 || M.dc > 0.0;
#endif
    M.it += 1;
    M.phase = BO_PH_EVAL;
    return -1;
  }
  // ---- not acceptable ----
  bool try_soc = false;
  if (M.soc == 0) {
    // second-order correction (Waechter & Biegler section 2.4) for the first, full trial step only
    if (M.ls == 0 && finite && thetat >= M.theta0 && (BO_ME + BO_MI) > 0) {
      BO_UNROLL
      for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_DX0, i) = SM(BO_OFF_DX, i);
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_RE, j) = M.a * SM(BO_OFF_CE, j) + SM(BO_OFF_CET, j);
      BO_NOUNROLL
      for (int i = 0; i < BO_MI; ++i) {
        const double s = SM(BO_OFF_S, i), ds = SM(BO_OFF_DS, i);
        SM(BO_OFF_DS0, i) = ds;
        SM(BO_OFF_RI, i) = M.a * (SM(BO_OFF_CI, i) - s) + (SM(BO_OFF_CIT, i) - (s + at * ds));
      }
      M.th_soc = thetat;
      try_soc = true;
    }
  } else if (M.soc < 4 && finite && thetat <= kappa_soc * M.th_soc) {
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_RE, j) = at * SM(BO_OFF_RE, j) + SM(BO_OFF_CET, j);
    BO_NOUNROLL
    for (int i = 0; i < BO_MI; ++i)
      SM(BO_OFF_RI, i) = at * SM(BO_OFF_RI, i) + (SM(BO_OFF_CIT, i) - (SM(BO_OFF_S, i) + at * SM(BO_OFF_DS, i)));
    M.th_soc = thetat;
    try_soc = true;
  }
  if (try_soc) {
    M.a_trial = bo_tm_step(sm, M.mu, M.rho, M.dc, M.tau, false);  // corrected direction; tried on the next trip
    M.soc += 1;
    return -1;
  }
  if (M.soc > 0) {  // corrections did not help: back to the uncorrected direction (and its residuals)
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_DX, i) = SM(BO_OFF_DX0, i);
    BO_NOUNROLL
    for (int i = 0; i < BO_MI; ++i) SM(BO_OFF_DS, i) = SM(BO_OFF_DS0, i);
    M.soc = 0;
  }
  M.a *= 0.5;
  M.a_trial = M.a;
  M.ls += 1;
  if (M.ls >= BO_LS_MAX || M.a < BO_ALPHA_MIN) {
    if (M.theta0 > 1e-7 * fmax(1.0, M.theta_min * 1e4) && M.n_resto < BO_RESTO_MAX_PHASES) {
      // no acceptable step at an infeasible point: feasibility restoration (the evaluation at x is still valid)
      M.n_resto += 1;
      M.resto = 1;
      M.thr0 = M.thr = bo_tm_resto_prepare(sm);
      M.dw = BO_RESTO_ZETA;
      M.dc = BO_RESTO_DC;
      M.attempt = 0;
      M.heavy = 0;
      M.phase = BO_PH_FACTOR;
      return -1;
    }
    // no acceptable step along this direction: convexify harder (see bo_ipm_reg.cuh)
    if (++M.heavy >= BO_HEAVY_MAX) return M.err0 <= prm.acceptable_tol ? BO_ST_ACCEPTABLE : BO_ST_LINE_SEARCH;
    M.dw = fmax(M.dw * 100.0, 1.0);
    M.phase = BO_PH_FACTOR;
  }
  return -1;
}

#ifdef BO_HOST_SIM
// Test harness: one instance, the roles executed one after the other where the kernel has a CTA barrier.
static int bo_team_solve_host(bo_tm& M, double* sm, const bo_solver_params& prm) {
  bo_tm_begin(M, sm, prm, 0);
  int status = -1;
  while (status < 0) {
    if (M.phase == BO_PH_EVAL)
      for (int r = 0; r < BO_G; ++r) bo_team_kkt(r, sm);
    if (M.phase != BO_PH_INIT) status = bo_tm_m1(M, sm, prm);
    if (status >= 0) break;
    SM(BO_OFF_AT, 0) = M.a_trial;
    if (M.phase == BO_PH_TRIAL || M.phase == BO_PH_INIT) {
      const double at = M.phase == BO_PH_INIT ? 0.0 : M.a_trial;
      for (int r = 0; r < BO_G; ++r) bo_team_pre(sm, at, r);
      for (int r = 0; r < BO_G; ++r) bo_team_fc(r, sm, at, M.phase == BO_PH_TRIAL);
    }
    status = bo_tm_m2(M, sm, prm);
  }
  return status;
}
#else

__device__ __forceinline__ unsigned bo_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bo_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bo_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bo_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bo_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bo_mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bo_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bo_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   bo_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(bo_smem_u32(bar))
               : "memory");
}

// Input staging: tiles of BO_TILE consecutive instances, rows of p and x0 back to back, fetched by the bulk-copy
// engine (TMA) one tile ahead of use.  Two buffers; `tile_b[s]` = first instance of the tile in buffer s (-1: none),
// `tile_n[s]` = instances in it, `tile_used[s]` = handed out so far.
#define BO_TILE 32
#define BO_STAGE_DOUBLES (BO_TILE * (BO_NP + BO_NX))
struct bo_stage_ctl {
  unsigned long long bar[2];
  long long tile_b[2];
  int tile_n[2], tile_used[2], parity[2], bulk[2];
  int cur, exhausted;
};

#ifndef BO_MIN_CTAS
#define BO_MIN_CTAS 2
#endif

extern "C" __global__ void __launch_bounds__(BO_TPB, BO_MIN_CTAS)
bo_solve_kernel(long long B, const double* __restrict__ p_all, const double* __restrict__ x0_all,
                double* __restrict__ x_all, double* __restrict__ lam_all, double* __restrict__ f_all,
                int* __restrict__ status_all, int* __restrict__ iters_all, double* __restrict__ kkt_all,
                unsigned long long* __restrict__ work_counter, const bo_solver_params prm) {
  extern __shared__ __align__(128) double bo_sm_all[];
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  double* const sm = bo_sm_all + lane;
  double* const stage = bo_sm_all + BO_SM_ELEMS * 32;                       // [2][BO_STAGE_DOUBLES]
  int* const ctrl = reinterpret_cast<int*>(stage + 2 * BO_STAGE_DOUBLES);   // [32] phase of every team
  bo_stage_ctl* const sc = reinterpret_cast<bo_stage_ctl*>(ctrl + 32);
  const bool aligned = ((((unsigned long long)p_all) | ((unsigned long long)x0_all)) & 15ULL) == 0ULL;

  bo_tm M;
  M.phase = BO_PH_IDLE;
  M.b = -1;

  // thread 0 of the master warp runs the staging pipeline
  auto issue_tile = [&](int s) {  // grab the next tile of instances and start its copies into buffer s
    const long long b0 = (long long)atomicAdd(work_counter, (unsigned long long)BO_TILE);
    if (b0 >= B) {
      sc->tile_b[s] = -1;
      sc->tile_n[s] = 0;
      sc->exhausted = 1;
      return;
    }
    const int n = (int)((B - b0) < (long long)BO_TILE ? (B - b0) : (long long)BO_TILE);
    sc->tile_b[s] = b0;
    sc->tile_n[s] = n;
    sc->tile_used[s] = 0;
    double* dst = stage + s * BO_STAGE_DOUBLES;
    // the previous tenant of this buffer was read through the generic proxy: order those reads before the engine's writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const bool bulk = aligned && n == BO_TILE;  // full tiles of 16-byte aligned rows go through the bulk-copy engine
    sc->bulk[s] = bulk ? 1 : 0;
    if (bulk) {
      unsigned bytes = 0;
      if (BO_NP > 0) bytes += (unsigned)(BO_TILE * BO_NP * sizeof(double));
      if (x0_all) bytes += (unsigned)(BO_TILE * BO_NX * sizeof(double));
      if (bytes > 0) {
        bo_mbar_expect_tx(&sc->bar[s], bytes);
        if (BO_NP > 0) bo_bulk_g2s(dst, p_all + b0 * BO_NP, (unsigned)(BO_TILE * BO_NP * sizeof(double)), &sc->bar[s]);
        if (x0_all) bo_bulk_g2s(dst + BO_TILE * BO_NP, x0_all + b0 * BO_NX, (unsigned)(BO_TILE * BO_NX * sizeof(double)), &sc->bar[s]);
      } else {
        sc->bulk[s] = 0;
      }
    }
  };

  if (threadIdx.x == 0) {
    bo_mbar_init(&sc->bar[0], 1);
    bo_mbar_init(&sc->bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sc->parity[0] = sc->parity[1] = 0;
    sc->cur = 0;
    sc->exhausted = 0;
    issue_tile(0);
    if (!sc->exhausted) issue_tile(1);
    else { sc->tile_b[1] = -1; sc->tile_n[1] = 0; }
  }
  if (role == 0) ctrl[lane] = BO_PH_IDLE;
  __syncthreads();

  // Master warp: give every idle team the next staged instance.  Warp-synchronous: lane 0 does the bookkeeping.
  auto fetch = [&]() {
    const unsigned idle = __ballot_sync(0xffffffffu, M.phase == BO_PH_IDLE);
    if (idle == 0u) return;
    int want = __popc(idle);
    const int my_rank = __popc(idle & ((1u << lane) - 1u));
    int given = 0;  // instances handed out so far in this call
    while (want > 0) {
      int s = 0, first = 0, take = 0;
      long long b0 = -1;
      if (lane == 0) {
        s = sc->cur;
        if (sc->tile_b[s] < 0 || sc->tile_used[s] >= sc->tile_n[s]) {
          // current buffer is spent: refill it (one tile ahead) and move on to the other one
          if (sc->tile_b[s] >= 0 || !sc->exhausted) {
            if (!sc->exhausted) issue_tile(s);
            else { sc->tile_b[s] = -1; sc->tile_n[s] = 0; }
          }
          s ^= 1;
          sc->cur = s;
        }
        if (sc->tile_b[s] >= 0 && sc->tile_used[s] < sc->tile_n[s]) {
          if (sc->tile_used[s] == 0 && sc->bulk[s]) {  // first use of this tile: its bytes must have landed
            bo_mbar_wait(&sc->bar[s], (unsigned)sc->parity[s]);
            sc->parity[s] ^= 1;
          }
          first = sc->tile_used[s];
          take = sc->tile_n[s] - first;
          if (take > want) take = want;
          sc->tile_used[s] = first + take;
          b0 = sc->tile_b[s];
        }
      }
      __syncwarp();
      s = __shfl_sync(0xffffffffu, s, 0);
      first = __shfl_sync(0xffffffffu, first, 0);
      take = __shfl_sync(0xffffffffu, take, 0);
      b0 = __shfl_sync(0xffffffffu, b0, 0);
      if (take == 0) break;  // nothing left anywhere
      if (M.phase == BO_PH_IDLE && my_rank >= given && my_rank < given + take) {
        const int k = first + (my_rank - given);  // my instance inside the tile
        const long long b = b0 + k;
        const int bulk = sc->bulk[s];
        const double* src_p = bulk ? stage + s * BO_STAGE_DOUBLES + k * BO_NP : p_all + b * BO_NP;
        const double* src_x = bulk ? stage + s * BO_STAGE_DOUBLES + BO_TILE * BO_NP + k * BO_NX : (x0_all ? x0_all + b * BO_NX : nullptr);
        BO_UNROLL
        for (int i = 0; i < BO_NP; ++i) SM(BO_OFF_P, i) = src_p[i];
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_X, i) = (x0_all != nullptr) ? src_x[i] : 0.0;
        bo_tm_begin(M, sm, prm, b);
      }
      given += take;
      want -= take;
      __syncwarp();
    }
  };

  auto finish = [&](int status) {  // master: results of a finished instance -> global memory
    const long long b = M.b;
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) x_all[b * BO_NX + i] = SM(BO_OFF_X, i);
    if (lam_all) {
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) lam_all[b * (BO_ME + BO_MI) + j] = SM(BO_OFF_Y, j);
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) lam_all[b * (BO_ME + BO_MI) + BO_ME + i] = SM(BO_OFF_Z, i);
    }
    if (f_all) f_all[b] = M.f;
    if (status_all) status_all[b] = status;
    if (iters_all) iters_all[b] = M.it;
    if (kkt_all) kkt_all[b] = M.err0;
    M.phase = BO_PH_IDLE;
  };

  while (true) {
    if (role == 0) {
      fetch();
      ctrl[lane] = M.phase;
    }
    if (!__syncthreads_or(role == 0 && M.phase != BO_PH_IDLE)) break;
    // ---- W1: KKT tape slices at x for the teams that start an iteration ----
    if (ctrl[lane] == BO_PH_EVAL) bo_team_kkt(role, sm);
    __syncthreads();
    // ---- M1 ----
    int status = -1;
    if (role == 0) {
      if (M.phase == BO_PH_EVAL || M.phase == BO_PH_FACTOR || M.phase == BO_PH_TRIAL) {
        status = bo_tm_m1(M, sm, prm);
        if (status >= 0) M.phase = BO_PH_DONE;
      }
      SM(BO_OFF_AT, 0) = M.a_trial;
      ctrl[lane] = M.phase;
    }
    __syncthreads();
    // ---- W2a: trial point (or the seed of a fresh instance) and the shared sin / cos of its components ----
    const int ph2 = ctrl[lane];
    const bool w2 = ph2 == BO_PH_TRIAL || ph2 == BO_PH_INIT;
    const double at2 = w2 ? SM(BO_OFF_AT, 0) : 0.0;
    if (w2) bo_team_pre(sm, at2, role);
    __syncthreads();
    // ---- W2b: f / c slices there, then the barrier work of the inequality rows each slice owns ----
    if (w2) bo_team_fc(role, sm, at2, ph2 == BO_PH_TRIAL);
    __syncthreads();
    // ---- M2 ----
    if (role == 0) {
      if (M.phase == BO_PH_TRIAL || M.phase == BO_PH_INIT) status = bo_tm_m2(M, sm, prm);
      if (status >= 0) finish(status);  // one copy of the write-out: instances that ended in M1 are parked in PH_DONE
    }
  }
}
#endif
