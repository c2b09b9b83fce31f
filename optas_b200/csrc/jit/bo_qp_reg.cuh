// bo_qp_reg.cuh -- convex-QP fast path of the thread-per-instance tiers (included by bo_ipm_reg.cuh when the
// generated prelude defines BO_QP).
//
// Role on the reference path: OSQPSolver / CVXOPTSolver (optas/solver.py:426-507, 514-580) and the qpOASES spelling of
// CasADiSolver for the QuadraticCost* problem classes (optas/optimization.py:219-260): P, q, M, c, A, b are functions
// of the parameters only, so the solvers there are handed constant matrices.  bo_problem_create decides it from the
// tape (no Jacobian / Hessian output depends on x, y or z: bo_codegen.cpp problem_is_qp), not from the class name.
//
// What that buys on the GPU:
//   * the tape runs ONCE per instance (at the seed): f, grad f, cE, cI are carried along every step exactly,
//       grad f += a H dx,   cE += a JE dx,   cI += a JI dx,   f += a g'dx + a^2/2 dx'H dx,
//     and once more at the end, so that what is reported (f, KKT error) is evaluated AT the x returned;
//   * no line search, no filter, no trial points: Mehrotra's predictor-corrector (the method CVXOPT's coneqp runs:
//     affine step -> centring parameter sigma = (mu_aff / mu)^3 -> corrected step, both from ONE factorisation of
//     H + JI' (Z/S) JI), one common primal / dual step length from the fraction-to-the-boundary rule;
//   * an iteration is one trip with the same control flow in every lane (factor, two substitutions).
// The linear algebra is the tier's own (rho-augmented unpivoted LDL', bo_ipm_reg.cuh), the optimality measure is the
// interior-point kernel's (scaled KKT error <= tol), so status / kkt_error mean the same on both paths.
#pragma once

#ifndef BO_QP_ALPHA_OK
#define BO_QP_ALPHA_OK 0.1  /* below this step length the corrected step is replaced by a centring step */
#endif
// Weight of the second-order term ds_aff * dz_aff in the corrector.  Mehrotra's method uses 1; when the affine step is
// cut short by the boundary (a_aff << 1) the full term describes a step that is never taken, and the plain method was
// seen cycling on badly centred iterates (16 of 65536 box QPs, mu alternating 0.008 <-> 0.025).  Weighting by a_aff
// removed every such case and lowers the mean iteration count (box QPs 7.37 -> 6.94).
#ifndef BO_QP_CROSS
#define BO_QP_CROSS(a) (a)
#endif
#ifndef BO_QP_IC_MAX
#define BO_QP_IC_MAX 12  /* regularisation attempts (semidefinite H / dependent rows of A) before giving up */
#endif

BO_NOINLINE void bo_qp_eval(bo_ipm_state& S) {
  // y, z do not enter any output of a QP's kkt tape other than through terms that vanish; pass the current ones
  bo_tape_kkt(S.x, S.p, S.y, S.z, &S.f, S.g, S.cE, S.cI, S.JE, S.JI, S.H);
}

BO_DEVICE void bo_ipm_init(bo_ipm_state& S, const bo_solver_params prm) {
  S.mu = prm.mu_init;
  S.it = 0;
  S.trips = 0;
  S.err0 = BO_INF;
  S.dw = 0.0;
  S.dc = 0.0;
  S.attempt = 0;
  S.ls = 0;
  S.soc = 1;  // "fresh": g, cE, cI, f were evaluated by the tape at the current x
  S.static_fac = true;
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) S.y[j] = 0.0;
  BO_UNROLL
  for (int i = 0; i < BO_MI; ++i) S.z[i] = 0.0;
  bo_qp_eval(S);
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) {
    S.s[i] = fmax(S.cI[i], 1e-2 * fmax(1.0, fabs(S.cI[i])));
    S.z[i] = S.mu / S.s[i];
  }
}

// Newton step for the complementarity target rc (given as rcs = rc / s): fills S.sol (dx, -dy), S.dx, S.ds and S.ds0 := dz.
BO_NOINLINE void bo_qp_step(bo_ipm_state& S, const bo_solver_params& prm, const double* BO_RESTRICT rcs) {
  double tvec[BO_DIM(BO_MI)];
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) tvec[i] = -(rcs[i] + S.sigma[i] * S.rI[i]);
  BO_UNROLL
  for (int i = 0; i < BO_NX; ++i) S.sol[i] = -S.rd[i];
  bo_JIt_acc(S.JI, tvec, S.sol);
  double t2[BO_DIM(BO_ME)];
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) {
    S.sol[BO_NX + j] = -S.cE[j];
    t2[j] = -S.rho * S.cE[j];
  }
  bo_JEt_acc(S.JE, t2, S.sol);
  BO_LDL_SOLVE(S.LD, S.sol);
  const double undo = 1.0 / (1.0 - S.rho * S.dc);
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) S.sol[BO_NX + j] *= undo;
  BO_UNROLL
  for (int i = 0; i < BO_NX; ++i) S.dx[i] = S.sol[i];
  bo_JI_mul(S.JI, S.dx, S.ds);
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) {
    S.ds[i] += S.rI[i];
    S.ds0[i] = -(rcs[i] + S.sigma[i] * S.ds[i]);  // dz
  }
}

// Largest a in (0, 1] with s + a ds >= (1 - tau) s and z + a dz >= (1 - tau) z.
BO_DEVICE double bo_qp_max_step(const bo_ipm_state& S, double tau) {
  double a = 1.0;
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) {
    if (S.ds[i] < 0.0) a = fmin(a, -tau * S.s[i] / S.ds[i]);
    if (S.ds0[i] < 0.0) a = fmin(a, -tau * S.z[i] / S.ds0[i]);
  }
  return a;
}

// One trip = one iteration: optimality test, factorisation, predictor, corrector, update.  Returns -1 to continue.
BO_DEVICE int bo_ipm_trip(bo_ipm_state& S, const bo_solver_params prm) {
  const double s_max = 100.0;
  const bool over = ++S.trips > prm.max_trips;
  double mu = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) S.rd[i] = S.g[i];
    {
      double ny[BO_DIM(BO_ME)], nz[BO_DIM(BO_MI)];
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) ny[j] = -S.y[j];
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) nz[i] = -S.z[i];
      bo_JEt_acc(S.JE, ny, S.rd);
      bo_JIt_acc(S.JI, nz, S.rd);
    }
    double e_dual = 0.0, e_prim = 0.0, e_comp = 0.0, sum_mult = 0.0, sum_z = 0.0;
    mu = 0.0;
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) e_dual = fmax(e_dual, fabs(S.rd[i]));
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) {
      e_prim = fmax(e_prim, fabs(S.cE[j]));
      sum_mult += fabs(S.y[j]);
    }
    BO_NOUNROLL
    for (int i = 0; i < BO_MI; ++i) {
      S.rI[i] = S.cI[i] - S.s[i];
      e_prim = fmax(e_prim, fabs(S.rI[i]));
      // complementarity against the slack AND against the constraint value itself (what a checker without slacks sees)
      e_comp = fmax(e_comp, fmax(S.s[i], fabs(S.cI[i])) * S.z[i]);
      mu += S.s[i] * S.z[i];
      sum_z += fabs(S.z[i]);
    }
    sum_mult += sum_z;
    mu /= (double)BO_DIM(BO_MI);
    const double s_d = (BO_ME + BO_MI) > 0 ? fmax(s_max, sum_mult / (double)BO_DIM(BO_ME + BO_MI)) / s_max : 1.0;
    const double s_c = BO_MI > 0 ? fmax(s_max, sum_z / (double)BO_DIM(BO_MI)) / s_max : 1.0;
    S.err0 = fmax(fmax(e_dual / s_d, e_prim), e_comp / s_c);
#ifdef BO_HOST_TRACE
    printf("qp it %3d f %.6e err0 %.3e (dual %.3e prim %.3e comp %.3e) mu %.2e fresh %d dw %.1e dc %.1e\n", S.it, S.f, S.err0,
           e_dual / s_d, e_prim, e_comp / s_c, mu, S.soc, S.dw, S.dc);
#endif
    if (!bo_isfinite(S.err0) || !bo_isfinite(S.f)) return BO_ST_NUMERICAL;
    const bool stop = S.err0 <= prm.tol || S.it >= prm.max_iter || over;
    if (!stop) break;
    if (S.soc) {
      if (S.err0 <= prm.tol) return BO_ST_CONVERGED;
      return S.err0 <= prm.acceptable_tol ? BO_ST_ACCEPTABLE : BO_ST_MAX_ITER;
    }
    bo_qp_eval(S);  // report (and decide on) values evaluated at the x that is returned, not carried along
    S.soc = 1;
  }
  S.mu = mu;

  // ---- factorisation of [H + JI' (Z/S) JI + rho JE'JE + dw, JE'; JE, -dc'] ----
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) S.sigma[i] = S.z[i] / S.s[i];
  bo_kkt_fill(S.H, S.JE, S.JI, S.sigma, BO_LDP(S));
  S.rho = BO_STATIC_RHO;
  bo_JEtJE_acc(S.JE, S.rho, BO_LDP(S));
#ifdef BO_SPARSE_LDL
  {
    const int32_t* sign = prm.ldl_tab + prm.ldl_tab[5];
    const double dcp = S.dc / (1.0 - S.rho * S.dc);
    for (int j = 0; j < BO_NK; ++j) BO_LDP(S)[(long long)j * BO_LDS(S)] += sign[j] > 0 ? S.dw : -dcp;
  }
  const int bad = bo_ldl_sparse(BO_LDP(S), BO_LDS(S), prm.ldl_tab);
#else
  for (int i = 0; i < BO_NX; ++i) S.LD[BO_KIDX(i, i)] += S.dw;
  for (int i = BO_NX; i < BO_NK; ++i) S.LD[BO_KIDX(i, i)] -= S.dc / (1.0 - S.rho * S.dc);
  const int bad = bo_ldl_static(S.LD);
#endif
  if (bad != 0) {
    // semidefinite Hessian with too few active rows, or linearly dependent equality rows: static regularisation, kept for
    // the rest of the instance (the matrix structure does not change); the steps then are inexact Newton steps towards
    // the same residuals, the optimality test above is unaffected
    if (bad == 2 && S.dc == 0.0) S.dc = 1e-9;
    else S.dw = S.dw == 0.0 ? 1e-8 : S.dw * 100.0;
    S.ls = 1;  // regularised in this iteration
    if (++S.attempt > BO_QP_IC_MAX) return S.err0 <= prm.acceptable_tol ? BO_ST_ACCEPTABLE : BO_ST_NUMERICAL;
    return -1;
  }

  // ---- predictor (affine scaling) and Mehrotra's corrector ----
  double rcs[BO_DIM(BO_MI)];
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) rcs[i] = S.z[i];
  bo_qp_step(S, prm, rcs);
  double alpha = 1.0;
  if (BO_MI > 0) {
    const double a_aff = bo_qp_max_step(S, 1.0);
    double mu_aff = 0.0;
    BO_NOUNROLL
    for (int i = 0; i < BO_MI; ++i) mu_aff += (S.s[i] + a_aff * S.ds[i]) * (S.z[i] + a_aff * S.ds0[i]);
    mu_aff /= (double)BO_DIM(BO_MI);
    const double r = mu > 0.0 ? mu_aff / mu : 0.0;
    // centring target sigma mu, not below tol / 10 (IPOPT's mu_min): driving the products s z further down only blows up
    // Z/S -- and with it the conditioning of the system -- while the dual residual is still converging
    const double smu = fmax(r * r * r * mu, fmin(mu, 0.1 * prm.tol));
    BO_NOUNROLL
    for (int i = 0; i < BO_MI; ++i) rcs[i] = S.z[i] + (BO_QP_CROSS(a_aff) * S.ds[i] * S.ds0[i] - smu) / S.s[i];
    bo_qp_step(S, prm, rcs);
    alpha = bo_qp_max_step(S, fmax(0.99, 1.0 - mu));
    if (alpha < BO_QP_ALPHA_OK) {
      // Badly centred iterate (a few products s z far above the mean): the corrected direction runs into the boundary at
      // once and the plain method can cycle.  Take a pure centring step instead (sigma = 1, no second-order term) with
      // the same factorisation; the next iteration then starts from balanced products.
      BO_NOUNROLL
      for (int i = 0; i < BO_MI; ++i) rcs[i] = S.z[i] - mu / S.s[i];
      bo_qp_step(S, prm, rcs);
      alpha = bo_qp_max_step(S, fmax(0.99, 1.0 - mu));
    }
#ifdef BO_HOST_TRACE
    printf("       a_aff %.3e sigma %.3e alpha %.3e\n", a_aff, r * r * r, alpha);
#endif
  }

  if (prm.max_step > 0.0) {  // an explicit step cap is honoured here as well (the default is off for problems without sin / cos)
    double dxn = 0.0;
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) dxn = fmax(dxn, fabs(S.dx[i]));
    if (alpha * dxn > prm.max_step) alpha = prm.max_step / dxn;
  }

  // ---- move, carrying the function values along (exact for a quadratic cost and linear constraints) ----
  {
    double hdx[BO_NX], t[BO_DIM(BO_ME > BO_MI ? BO_ME : BO_MI)];
    bo_H_mul(S.H, S.dx, hdx);
    double gdx = 0.0, dhd = 0.0;
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) {
      gdx += S.g[i] * S.dx[i];
      dhd += hdx[i] * S.dx[i];
    }
    S.f += alpha * gdx + 0.5 * alpha * alpha * dhd;
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) {
      S.g[i] += alpha * hdx[i];
      S.x[i] += alpha * S.dx[i];
    }
    if (BO_ME > 0) {
      bo_JE_mul(S.JE, S.dx, t);
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) {
        S.cE[j] += alpha * t[j];
        S.y[j] -= alpha * S.sol[BO_NX + j];
      }
    }
    if (BO_MI > 0) {
      bo_JI_mul(S.JI, S.dx, t);
      BO_NOUNROLL
      for (int i = 0; i < BO_MI; ++i) {
        S.cI[i] += alpha * t[i];
        S.s[i] += alpha * S.ds[i];
        S.z[i] += alpha * S.ds0[i];
      }
    }
  }
  S.soc = 0;
  ++S.it;
  // a regularisation that was needed once is tried smaller the next time (it is not a property of the instance: the
  // conditioning of H + JI' (Z/S) JI changes with Z/S)
  if (S.ls) S.dw *= 1e-2;
  if (S.dw < 1e-8) S.dw = 0.0;
  S.ls = 0;
  S.attempt = 0;
  return -1;
}

// Convenience driver for one instance (used by the host-compiled test harness).
BO_DEVICE int bo_ipm_solve(bo_ipm_state& S, const bo_solver_params prm) {
  bo_ipm_init(S, prm);
  int status;
  do {
    status = bo_ipm_trip(S, prm);
  } while (status < 0);
  return status;
}

#ifndef BO_HOST_SIM
// Persistent lanes with per-lane work fetching, as the interior-point kernel of this tier (bo_ipm_reg.cuh); a trip is
// one short block of code here, so there is no barrier between phases.
extern "C" __global__ void __launch_bounds__(BO_TPB)
bo_solve_kernel(long long B, const double* __restrict__ p_all, const double* __restrict__ x0_all,
                double* __restrict__ x_all, double* __restrict__ lam_all, double* __restrict__ f_all,
                int* __restrict__ status_all, int* __restrict__ iters_all, double* __restrict__ kkt_all,
                unsigned long long* __restrict__ work_counter, const bo_solver_params prm) {
  bo_ipm_state S;
  long long b = -1;
  bool active = false, exhausted = false;
  while (true) {
    if (!active && !exhausted) {
      b = (long long)atomicAdd(work_counter, 1ULL);
      if (b < B) {
        BO_UNROLL
        for (int i = 0; i < BO_NP; ++i) S.p[i] = p_all[b * BO_NP + i];
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) S.x[i] = x0_all ? x0_all[b * BO_NX + i] : 0.0;
        bo_ipm_init(S, prm);
        active = true;
      } else {
        exhausted = true;
      }
    }
    if (!__any_sync(0xffffffffu, active)) break;
    if (active) {
      const int status = bo_ipm_trip(S, prm);
      if (status >= 0) {
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) x_all[b * BO_NX + i] = S.x[i];
        if (lam_all) {
          BO_UNROLL
          for (int j = 0; j < BO_ME; ++j) lam_all[b * (BO_ME + BO_MI) + j] = S.y[j];
          BO_UNROLL
          for (int i = 0; i < BO_MI; ++i) lam_all[b * (BO_ME + BO_MI) + BO_ME + i] = S.z[i];
        }
        if (f_all) f_all[b] = S.f;
        if (status_all) status_all[b] = status;
        if (iters_all) iters_all[b] = S.it;
        if (kkt_all) kkt_all[b] = S.err0;
        active = false;
      }
    }
  }
}
#endif
