// bo_ipm_reg.cuh -- register-resident primal-dual interior-point / Newton-KKT solver,
// ONE PROBLEM INSTANCE PER THREAD (tier "S": nx + n_eq up to a few tens).
//
// Replaces, for a whole batch at once, what the reference does per call inside
// CasADiSolver._solve (optas/solver.py:386-398 -> casadi nlpsol("ipopt")): evaluate
// f, grad f, c, Jacobians and the Hessian of the Lagrangian (here: straight-line code generated
// from the problem's expression tapes, see bo_codegen.cpp), assemble the primal-dual KKT system,
// factor it (dense LDL' with inertia-correcting regularisation), line-search, update the barrier.
//
// This header is included AFTER the generated prelude, which defines
//   BO_NX, BO_NP, BO_ME, BO_MI, BO_NNZ_JE, BO_NNZ_JI, BO_NNZ_H, BO_TPB
//   bo_tape_fc(x, p, f, cE, cI)
//   bo_tape_kkt(x, p, y, z, f, g, cE, cI, JE, JI, H)
//   bo_JEt_acc / bo_JIt_acc (out += J' v), bo_JE_mul / bo_JI_mul (out = J v),
//   bo_kkt_fill(H, JE, JI, sigma, K), bo_xHx(H, v)
//
// Problem:  min f(x)  s.t.  cE(x) = 0,  cI(x) - s = 0,  s >= 0      (s: slacks)
// Lagrangian L = f - y'cE - z'cI,  z >= 0.  Barrier sub-problem parameter mu.
// Reduced Newton system solved each iteration (ds, dz eliminated):
//   [ H + JI' S JI + dw I    JE'   ] [ dx  ]   [ -(grad L) - JI' ((z*cI - mu)/s) ]
//   [ JE                    -dc I  ] [ -dy ] = [ -cE                            ],   S = diag(z/s)
#pragma once
#include "bo_common.cuh"

#define BO_NK (BO_NX + BO_ME)
#define BO_KSZ (BO_NK * BO_NK)
#define BO_KIDX(i, j) ((i) * BO_NK + (j)) /* row-major, lower triangle (i >= j) is the data */
#define BO_DIM(n) ((n) > 0 ? (n) : 1)

// Bunch-Kaufman LDL' (diagonal pivoting with 1x1 and 2x2 blocks; the unblocked LAPACK dsytf2
// algorithm, lower variant) of the symmetric indefinite KKT matrix, in place.  Pivoting is data
// dependent, so A lives in thread-local memory (L1-resident: 800 B for the 7-DoF IK problem).
// Outputs the pivot record ipiv (LAPACK convention, 1-based, negative for 2x2 blocks) and the
// inertia; returns false if a pivot block is numerically singular.
BO_DEVICE bool bo_bk_factor(double* BO_RESTRICT A, int* BO_RESTRICT ipiv, int* n_neg_out) {
  const double alpha = 0.6403882032022076;  // (1 + sqrt(17)) / 8
  // Like LAPACK, only an exactly (here: denormal-small) zero pivot is "singular"; near-singular
  // systems show up as a wrong inertia count and are handled by the caller's regularisation.
  const double tiny = 1e-250;
  int n_neg = 0;
  bool ok = true;
  int k = 0;
  while (k < BO_NK) {
    int kstep = 1, kp = k, imax = k;
    const double absakk = fabs(A[BO_KIDX(k, k)]);
    double colmax = 0.0;
    for (int i = k + 1; i < BO_NK; ++i) {
      const double v = fabs(A[BO_KIDX(i, k)]);
      if (v > colmax) { colmax = v; imax = i; }
    }
    if (fmax(absakk, colmax) <= tiny) {
      ok = false;  // singular pivot column: caller regularises
      ipiv[k] = k + 1;
      ++k;
      continue;
    }
    if (absakk < alpha * colmax) {
      double rowmax = 0.0;
      for (int j = k; j < imax; ++j) rowmax = fmax(rowmax, fabs(A[BO_KIDX(imax, j)]));
      for (int i = imax + 1; i < BO_NK; ++i) rowmax = fmax(rowmax, fabs(A[BO_KIDX(i, imax)]));
      if (absakk >= alpha * colmax * (colmax / rowmax)) {
        kp = k;
      } else if (fabs(A[BO_KIDX(imax, imax)]) >= alpha * rowmax) {
        kp = imax;
      } else {
        kp = imax;
        kstep = 2;
      }
    }
    const int kk = k + kstep - 1;
    if (kp != kk) {  // symmetric interchange of rows/columns kk and kp in the trailing block
      for (int i = kp + 1; i < BO_NK; ++i) {
        const double tmp = A[BO_KIDX(i, kk)];
        A[BO_KIDX(i, kk)] = A[BO_KIDX(i, kp)];
        A[BO_KIDX(i, kp)] = tmp;
      }
      for (int j = kk + 1; j < kp; ++j) {
        const double tmp = A[BO_KIDX(j, kk)];
        A[BO_KIDX(j, kk)] = A[BO_KIDX(kp, j)];
        A[BO_KIDX(kp, j)] = tmp;
      }
      {
        const double tmp = A[BO_KIDX(kk, kk)];
        A[BO_KIDX(kk, kk)] = A[BO_KIDX(kp, kp)];
        A[BO_KIDX(kp, kp)] = tmp;
      }
      if (kstep == 2) {
        const double tmp = A[BO_KIDX(k + 1, k)];
        A[BO_KIDX(k + 1, k)] = A[BO_KIDX(kp, k)];
        A[BO_KIDX(kp, k)] = tmp;
      }
    }
    if (kstep == 1) {
      const double akk = A[BO_KIDX(k, k)];
      if (fabs(akk) <= tiny) ok = false;
      if (akk < 0.0) ++n_neg;
      const double r1 = 1.0 / akk;
      for (int j = k + 1; j < BO_NK; ++j) {
        const double w = r1 * A[BO_KIDX(j, k)];
        for (int i = j; i < BO_NK; ++i) A[BO_KIDX(i, j)] -= A[BO_KIDX(i, k)] * w;
      }
      for (int i = k + 1; i < BO_NK; ++i) A[BO_KIDX(i, k)] *= r1;
      ipiv[k] = kp + 1;
    } else {
      ++n_neg;  // a 2x2 pivot block has one positive and one negative eigenvalue
      const double a21 = A[BO_KIDX(k + 1, k)];
      const double d11 = A[BO_KIDX(k + 1, k + 1)] / a21, d22 = A[BO_KIDX(k, k)] / a21;
      const double tt = 1.0 / (d11 * d22 - 1.0), d21 = tt / a21;
      for (int j = k + 2; j < BO_NK; ++j) {
        const double wk = d21 * (d11 * A[BO_KIDX(j, k)] - A[BO_KIDX(j, k + 1)]);
        const double wkp1 = d21 * (d22 * A[BO_KIDX(j, k + 1)] - A[BO_KIDX(j, k)]);
        for (int i = j; i < BO_NK; ++i) A[BO_KIDX(i, j)] -= A[BO_KIDX(i, k)] * wk + A[BO_KIDX(i, k + 1)] * wkp1;
        A[BO_KIDX(j, k)] = wk;
        A[BO_KIDX(j, k + 1)] = wkp1;
      }
      ipiv[k] = -(kp + 1);
      ipiv[k + 1] = -(kp + 1);
    }
    k += kstep;
  }
  *n_neg_out = n_neg;
  return ok;
}

// Solve A x = b with the factorisation above (LAPACK dsytrs, lower variant), in place.
BO_DEVICE void bo_bk_solve(const double* BO_RESTRICT A, const int* BO_RESTRICT ipiv, double* BO_RESTRICT b) {
  int k = 0;
  while (k < BO_NK) {  // forward: L D
    if (ipiv[k] > 0) {
      const int kp = ipiv[k] - 1;
      const double tmp = b[k]; b[k] = b[kp]; b[kp] = tmp;
      for (int i = k + 1; i < BO_NK; ++i) b[i] -= A[BO_KIDX(i, k)] * b[k];
      b[k] /= A[BO_KIDX(k, k)];
      k += 1;
    } else {
      const int kp = -ipiv[k] - 1;
      const double tmp = b[k + 1]; b[k + 1] = b[kp]; b[kp] = tmp;
      for (int i = k + 2; i < BO_NK; ++i) b[i] -= A[BO_KIDX(i, k)] * b[k] + A[BO_KIDX(i, k + 1)] * b[k + 1];
      const double akm1k = A[BO_KIDX(k + 1, k)];
      const double akm1 = A[BO_KIDX(k, k)] / akm1k, ak = A[BO_KIDX(k + 1, k + 1)] / akm1k;
      const double denom = akm1 * ak - 1.0;
      const double bkm1 = b[k] / akm1k, bk = b[k + 1] / akm1k;
      b[k] = (ak * bkm1 - bk) / denom;
      b[k + 1] = (akm1 * bk - bkm1) / denom;
      k += 2;
    }
  }
  k = BO_NK - 1;
  while (k >= 0) {  // backward: L'
    if (ipiv[k] > 0) {
      for (int i = k + 1; i < BO_NK; ++i) b[k] -= A[BO_KIDX(i, k)] * b[i];
      const int kp = ipiv[k] - 1;
      const double tmp = b[k]; b[k] = b[kp]; b[kp] = tmp;
      k -= 1;
    } else {
      for (int i = k + 1; i < BO_NK; ++i) {
        b[k] -= A[BO_KIDX(i, k)] * b[i];
        b[k - 1] -= A[BO_KIDX(i, k - 1)] * b[i];
      }
      const int kp = -ipiv[k] - 1;
      const double tmp = b[k]; b[k] = b[kp]; b[kp] = tmp;
      k -= 2;
    }
  }
}

// Factor K + diag(dw I, -dc I).  Returns 0 when the inertia is (BO_NX, BO_ME, 0), +1 when there
// are too many negative eigenvalues (reduced Hessian not positive definite), -1 when singular.
BO_DEVICE int bo_kkt_factor(const double* BO_RESTRICT K, double dw, double dc, double* BO_RESTRICT LD, int* BO_RESTRICT ipiv) {
  for (int i = 0; i < BO_NK; ++i)
    for (int j = 0; j <= i; ++j) LD[BO_KIDX(i, j)] = K[BO_KIDX(i, j)];
  for (int i = 0; i < BO_NX; ++i) LD[BO_KIDX(i, i)] += dw;
  for (int i = BO_NX; i < BO_NK; ++i) LD[BO_KIDX(i, i)] -= dc;
  int n_neg = 0;
  const bool okf = bo_bk_factor(LD, ipiv, &n_neg);
#ifdef BO_HOST_TRACE
  if (!okf || n_neg != BO_ME) {
    printf("       bk ok %d n_neg %d; D:", (int)okf, n_neg);
    for (int i = 0; i < BO_NK; ++i) printf(" %.2e(%d)", LD[BO_KIDX(i, i)], ipiv[i]);
    printf("\n");
  }
#endif
  if (!okf) return -1;
  if (n_neg == BO_ME) return 0;
  return n_neg > BO_ME ? 1 : -1;
}

// Barrier objective and l1 constraint violation at (x, s) given the function values there.
BO_DEVICE void bo_measures(double f, const double* cE, const double* cI, const double* s, double mu, double* phi,
                           double* theta) {
  double viol = 0.0, bar = 0.0;
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) viol += fabs(cE[j]);
  BO_UNROLL
  for (int i = 0; i < BO_MI; ++i) {
    viol += fabs(cI[i] - s[i]);
    bar += log(s[i]);
  }
  *theta = viol;
  *phi = f - mu * bar;
}

#define BO_NFILTER 8
#define BO_LS_MAX 16    /* backtracking halvings per direction (alpha >= 1.5e-5 alpha_max) */
#define BO_HEAVY_MAX 5  /* re-solves with dw = 1, 1e2, 1e4, 1e6 when no step is acceptable */

// Everything an instance carries from one iteration to the next.  A GPU lane owns one of these and
// re-uses it for instance after instance (see bo_solve_kernel).
struct bo_ipm_state {
  double p[BO_DIM(BO_NP)], x[BO_NX], s[BO_DIM(BO_MI)], y[BO_DIM(BO_ME)], z[BO_DIM(BO_MI)];
  double fth[BO_NFILTER], fph[BO_NFILTER];
  double f, mu, dw_last, err0, theta_max, theta_min;
  int nf, it, n_acceptable;
  bool recalc_y;
};

// Start an instance: S.p and S.x hold the parameters and the seed.
BO_DEVICE void bo_ipm_init(bo_ipm_state& S, const bo_solver_params prm) {
  double cE[BO_DIM(BO_ME)], cI[BO_DIM(BO_MI)];
  S.mu = prm.mu_init;
  S.dw_last = 0.0;
  S.err0 = BO_INF;
  S.theta_max = BO_INF;
  S.theta_min = 0.0;
  S.nf = 0;
  S.it = 0;
  S.n_acceptable = 0;
  S.recalc_y = false;
  bo_tape_fc(S.x, S.p, &S.f, cE, cI);
  BO_UNROLL
  for (int i = 0; i < BO_MI; ++i) {
    S.s[i] = fmax(cI[i], 1e-2 * fmax(1.0, fabs(cI[i])));
    S.z[i] = S.mu / S.s[i];
  }
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) S.y[j] = 0.0;
}

// One interior-point iteration.  Returns -1 to continue, or the final BO_ST_* status.
//
// Globalisation is IPOPT's filter line search (Waechter & Biegler 2006, section 2.3, with their
// default constants) on the pair (theta = ||c||_1, phi = barrier objective), plus a second-order
// correction for the first trial step.  IPOPT's restoration phase is replaced by re-solving the
// step with a heavily convexified Hessian (dw -> large turns the step into the minimum-norm
// feasibility step), accepted on constraint-violation decrease.
BO_DEVICE int bo_ipm_iterate(bo_ipm_state& S, const bo_solver_params prm) {
  const double kappa_eps = 10.0, kappa_mu = 0.2, theta_mu = 1.5, tau_min = 0.99, s_max = 100.0;
  const double kappa_sigma = 1e10, gamma_theta = 1e-5, gamma_phi = 1e-5, eta_phi = 1e-8;
  const double s_phi = 2.3, s_theta = 1.1, kappa_soc = 0.99;
  const double mu_min = prm.tol * 0.1;

  double *const p = S.p, *const x = S.x, *const s = S.s, *const y = S.y, *const z = S.z;
  double *const fth = S.fth, *const fph = S.fph;
  double &f = S.f, &mu = S.mu, &dw_last = S.dw_last, &err0 = S.err0, &theta_max = S.theta_max, &theta_min = S.theta_min;
  int &nf = S.nf, &n_acceptable = S.n_acceptable;
  bool& recalc_y = S.recalc_y;
  const int it = S.it;

  double cE[BO_DIM(BO_ME)], cI[BO_DIM(BO_MI)];
  double g[BO_NX], JE[BO_DIM(BO_NNZ_JE)], JI[BO_DIM(BO_NNZ_JI)], H[BO_DIM(BO_NNZ_H)];
  double K[BO_KSZ], LD[BO_KSZ], sol[BO_NK];
  int ipiv[BO_NK];
  double dx[BO_NX], ds[BO_DIM(BO_MI)], dz[BO_DIM(BO_MI)], sigma[BO_DIM(BO_MI)];
  double xt[BO_NX], st[BO_DIM(BO_MI)], cEt[BO_DIM(BO_ME)], cIt[BO_DIM(BO_MI)], rd[BO_NX];
  double rE[BO_DIM(BO_ME)], rI[BO_DIM(BO_MI)];  // constraint residuals the step is asked to remove
  {
    bo_tape_kkt(x, p, y, z, &f, g, cE, cI, JE, JI, H);
    if (recalc_y && BO_ME > 0) {
      // The last step needed Hessian convexification (dw > 0): its Newton multipliers scale with dw
      // and feed back into the Hessian.  Replace y by the least-squares estimate
      //   [ I  JE' ; JE  -dc ] [ r ; y ] = [ grad f - JI' z ; 0 ]
      // and re-evaluate the Hessian with it.
      recalc_y = false;
      BO_UNROLL
      for (int i = 0; i < BO_DIM(BO_NNZ_H); ++i) H[i] = 0.0;
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) sigma[i] = 0.0;
      bo_kkt_fill(H, JE, JI, sigma, K);
      if (bo_kkt_factor(K, 1.0, 1e-10, LD, ipiv) == 0) {
        double nz[BO_DIM(BO_MI)];
        BO_UNROLL
        for (int i = 0; i < BO_MI; ++i) nz[i] = -z[i];
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) sol[i] = g[i];
        bo_JIt_acc(JI, nz, sol);
        BO_UNROLL
        for (int j = 0; j < BO_ME; ++j) sol[BO_NX + j] = 0.0;
        bo_bk_solve(LD, ipiv, sol);
        bool fin = true;
        BO_UNROLL
        for (int j = 0; j < BO_ME; ++j) fin = fin && bo_isfinite(sol[BO_NX + j]);
        if (fin) {
          BO_UNROLL
          for (int j = 0; j < BO_ME; ++j) y[j] = sol[BO_NX + j];
        }
      }
      bo_tape_kkt(x, p, y, z, &f, g, cE, cI, JE, JI, H);
    }

    // ---- residuals and the scaled optimality error (IPOPT's E_mu, Waechter & Biegler eq. 5) ----
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) rd[i] = g[i];
    {
      double ny[BO_DIM(BO_ME)], nz[BO_DIM(BO_MI)];
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) ny[j] = -y[j];
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) nz[i] = -z[i];
      bo_JEt_acc(JE, ny, rd);
      bo_JIt_acc(JI, nz, rd);
    }
    double e_dual = 0.0, e_prim = 0.0, e_comp0 = 0.0, sum_mult = 0.0, sum_z = 0.0;
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) e_dual = fmax(e_dual, fabs(rd[i]));
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) {
      e_prim = fmax(e_prim, fabs(cE[j]));
      sum_mult += fabs(y[j]);
    }
    BO_UNROLL
    for (int i = 0; i < BO_MI; ++i) {
      e_prim = fmax(e_prim, fabs(cI[i] - s[i]));
      e_comp0 = fmax(e_comp0, s[i] * z[i]);
      sum_z += fabs(z[i]);
    }
    sum_mult += sum_z;
    const double s_d = (BO_ME + BO_MI) > 0 ? fmax(s_max, sum_mult / (double)BO_DIM(BO_ME + BO_MI)) / s_max : 1.0;
    const double s_c = BO_MI > 0 ? fmax(s_max, sum_z / (double)BO_DIM(BO_MI)) / s_max : 1.0;
    err0 = fmax(fmax(e_dual / s_d, e_prim), e_comp0 / s_c);
#ifdef BO_HOST_TRACE
    printf("it %3d f %.6e err0 %.3e (dual %.3e prim %.3e comp %.3e) mu %.2e nf %d dw_last %.2e\n", it, f, err0, e_dual / s_d,
           e_prim, e_comp0 / s_c, mu, nf, dw_last);
#endif
    if (!bo_isfinite(err0) || !bo_isfinite(f)) return BO_ST_NUMERICAL;
    if (err0 <= prm.tol) return BO_ST_CONVERGED;
    n_acceptable = (err0 <= prm.acceptable_tol) ? n_acceptable + 1 : 0;
    if (n_acceptable >= 15) return BO_ST_ACCEPTABLE;
    if (it >= prm.max_iter) return BO_ST_MAX_ITER;

    // ---- barrier parameter update (monotone Fiacco-McCormick, IPOPT eq. 7); resets the filter ----
    if (BO_MI > 0) {
      for (int rep = 0; rep < 8; ++rep) {
        double e_comp = 0.0;
        BO_UNROLL
        for (int i = 0; i < BO_MI; ++i) e_comp = fmax(e_comp, fabs(s[i] * z[i] - mu));
        const double err_mu = fmax(fmax(e_dual / s_d, e_prim), e_comp / s_c);
        if (err_mu <= kappa_eps * mu && mu > mu_min) {
          mu = fmax(mu_min, fmin(kappa_mu * mu, pow(mu, theta_mu)));
          nf = 0;
        } else {
          break;
        }
      }
    }
    const double tau = fmax(tau_min, 1.0 - mu);

    // ---- assemble and factor the reduced KKT system, with inertia correction (IPOPT Alg. IC) ----
    BO_UNROLL
    for (int i = 0; i < BO_MI; ++i) sigma[i] = z[i] / s[i];
    bo_kkt_fill(H, JE, JI, sigma, K);
    double dw = 0.0, dc = 0.0;
    int inertia = 1;
    for (int attempt = 0; attempt < 60; ++attempt) {
      inertia = bo_kkt_factor(K, dw, dc, LD, ipiv);
#ifdef BO_HOST_TRACE
      if (inertia != 0) printf("     inertia %d at dw %.3e dc %.3e\n", inertia, dw, dc);
#endif
      if (inertia == 0) break;
      if (inertia < 0 && BO_ME > 0 && dc == 0.0) {
        dc = 1e-8 * pow(mu, 0.25);  // singular: perturb the constraint block first
        continue;
      }
      if (dw == 0.0) {
        dw = (dw_last == 0.0) ? 1e-4 : fmax(1e-20, dw_last / 3.0);
      } else {
        dw *= (dw_last == 0.0) ? 100.0 : 8.0;
      }
      if (dw > 1e40) break;
    }
    if (inertia != 0) return BO_ST_NUMERICAL;
    if (dw > 0.0) dw_last = dw;

    // current measures, filter thresholds
    double phi0, theta0;
    bo_measures(f, cE, cI, s, mu, &phi0, &theta0);
    if (it == 0) {
      theta_max = 1e4 * fmax(1.0, theta0);
      theta_min = 1e-4 * fmax(1.0, theta0);
    }

    // step for given constraint residuals (rE, rI): fills sol (dx, -dy), dx, ds and returns the
    // fraction-to-the-boundary primal step length
    auto compute_step = [&]() -> double {
      double tvec[BO_DIM(BO_MI)];
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) tvec[i] = -(z[i] - mu / s[i] + sigma[i] * rI[i]);
      BO_UNROLL
      for (int i = 0; i < BO_NX; ++i) sol[i] = -rd[i];
      bo_JIt_acc(JI, tvec, sol);
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) sol[BO_NX + j] = -rE[j];
      bo_bk_solve(LD, ipiv, sol);
      BO_UNROLL
      for (int i = 0; i < BO_NX; ++i) dx[i] = sol[i];
      bo_JI_mul(JI, dx, ds);
      double ap = 1.0;
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) {
        ds[i] += rI[i];
        if (ds[i] < 0.0) ap = fmin(ap, -tau * s[i] / ds[i]);
      }
      return ap;
    };
    auto filter_ok = [&](double th, double ph) -> bool {
      if (!(th <= theta_max)) return false;
      for (int j = 0; j < nf; ++j)
        if (!(th <= (1.0 - gamma_theta) * fth[j] || ph <= fph[j] - gamma_phi * fth[j])) return false;
      return true;
    };

    bool accepted = false;
    double a_used = 0.0, y_step[BO_DIM(BO_ME)];
    for (int heavy = 0; heavy < BO_HEAVY_MAX && !accepted; ++heavy) {
      if (heavy > 0) {
        // no acceptable step: convexify harder (dw large => minimum-norm feasibility step); this
        // stands in for IPOPT's restoration phase on these small problems
        dw = fmax(dw * 100.0, 1.0);
        if (bo_kkt_factor(K, dw, dc, LD, ipiv) != 0) continue;
      }
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) rE[j] = cE[j];
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) rI[i] = cI[i] - s[i];
      const double a_p = compute_step();
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) dz[i] = -z[i] + mu / s[i] - sigma[i] * ds[i];
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) y_step[j] = -sol[BO_NX + j];
      double dphi = 0.0;  // directional derivative of the barrier objective
      BO_UNROLL
      for (int i = 0; i < BO_NX; ++i) dphi += g[i] * dx[i];
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) dphi -= mu * ds[i] / s[i];

      double a = a_p;
      if (prm.max_step > 0.0) {
        // step-length cap (cf. SNOPT's "major step limit"): Newton steps of several radians through
        // trigonometric kinematics are meaningless and wreck the multipliers
        double dxn = 0.0;
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) dxn = fmax(dxn, fabs(dx[i]));
        if (a * dxn > prm.max_step) a = prm.max_step / dxn;
      }
      for (int ls = 0; ls < BO_LS_MAX && !accepted; ++ls) {
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) xt[i] = x[i] + a * dx[i];
        BO_UNROLL
        for (int i = 0; i < BO_MI; ++i) st[i] = s[i] + a * ds[i];
        double ft, phit, thetat;
        bo_tape_fc(xt, p, &ft, cEt, cIt);
        bo_measures(ft, cEt, cIt, st, mu, &phit, &thetat);
        const bool finite = bo_isfinite(phit) && bo_isfinite(thetat);
        const bool ftype = dphi < 0.0 && a * pow(-dphi, s_phi) > pow(theta0, s_theta) && theta0 <= theta_min;
        const double slack = 10.0 * 2.2e-16 * fabs(phi0);
        bool ok = false, armijo = false;
        if (finite && filter_ok(thetat, phit)) {
          if (ftype) {
            armijo = phit - phi0 - slack <= eta_phi * a * dphi;
            ok = armijo;
          } else {
            ok = thetat <= (1.0 - gamma_theta) * theta0 || phit - slack <= phi0 - gamma_phi * theta0;
          }
        }
#ifdef BO_HOST_TRACE
        if (ls == 0 || ok)
          printf("     heavy %d ls %d a %.3e ok %d ftype %d theta %.3e->%.3e phi %.8e->%.8e dphi %.3e dw %.2e\n", heavy, ls, a,
                 (int)ok, (int)ftype, theta0, thetat, phi0, phit, dphi, dw);
#endif
        if (!ok && ls == 0 && finite && thetat >= theta0 && (BO_ME + BO_MI) > 0) {
          // ---- second-order correction (Waechter & Biegler section 2.4) ----
          double theta_old = theta0, th_soc = thetat;
          double dx0[BO_NX], ds0[BO_DIM(BO_MI)];
          BO_UNROLL
          for (int i = 0; i < BO_NX; ++i) dx0[i] = dx[i];
          BO_UNROLL
          for (int i = 0; i < BO_MI; ++i) ds0[i] = ds[i];
          BO_UNROLL
          for (int j = 0; j < BO_ME; ++j) rE[j] = a * cE[j] + cEt[j];
          BO_UNROLL
          for (int i = 0; i < BO_MI; ++i) rI[i] = a * (cI[i] - s[i]) + (cIt[i] - st[i]);
          for (int soc = 0; soc < 4; ++soc) {
            const double a_soc = compute_step();
            BO_UNROLL
            for (int i = 0; i < BO_NX; ++i) xt[i] = x[i] + a_soc * dx[i];
            BO_UNROLL
            for (int i = 0; i < BO_MI; ++i) st[i] = s[i] + a_soc * ds[i];
            bo_tape_fc(xt, p, &ft, cEt, cIt);
            bo_measures(ft, cEt, cIt, st, mu, &phit, &thetat);
            bool ok_soc = false;
            if (bo_isfinite(phit) && bo_isfinite(thetat) && filter_ok(thetat, phit)) {
              if (ftype) {
                armijo = phit - phi0 - slack <= eta_phi * a * dphi;
                ok_soc = armijo;
              } else {
                ok_soc = thetat <= (1.0 - gamma_theta) * theta0 || phit - slack <= phi0 - gamma_phi * theta0;
              }
            }
#ifdef BO_HOST_TRACE
            printf("       soc %d a %.3e ok %d theta %.3e phi %.8e\n", soc, a_soc, (int)ok_soc, thetat, phit);
#endif
            if (ok_soc) {
              ok = true;
              BO_UNROLL
              for (int i = 0; i < BO_MI; ++i) dz[i] = -z[i] + mu / s[i] - sigma[i] * ds[i];
              break;
            }
            if (!(thetat <= kappa_soc * th_soc)) break;
            th_soc = thetat;
            (void)theta_old;
            BO_UNROLL
            for (int j = 0; j < BO_ME; ++j) rE[j] = a_soc * rE[j] + cEt[j];
            BO_UNROLL
            for (int i = 0; i < BO_MI; ++i) rI[i] = a_soc * rI[i] + (cIt[i] - st[i]);
          }
          if (!ok) {  // restore the uncorrected direction for the backtracking that follows
            BO_UNROLL
            for (int i = 0; i < BO_NX; ++i) dx[i] = dx0[i];
            BO_UNROLL
            for (int i = 0; i < BO_MI; ++i) ds[i] = ds0[i];
          }
        }
        if (ok) {
          accepted = true;
          a_used = a;
          if (!(ftype && armijo) && nf < BO_NFILTER) {  // augment the filter (eq. 22)
            fth[nf] = (1.0 - gamma_theta) * theta0;
            fph[nf] = phi0 - gamma_phi * theta0;
            ++nf;
          } else if (!(ftype && armijo)) {
            int worst = 0;  // filter full: overwrite the entry with the largest theta
            for (int j = 1; j < BO_NFILTER; ++j)
              if (fth[j] > fth[worst]) worst = j;
            fth[worst] = (1.0 - gamma_theta) * theta0;
            fph[worst] = phi0 - gamma_phi * theta0;
          }
          break;
        }
        a *= 0.5;
        if (a < 1e-12) break;
      }
    }
    if (!accepted) return BO_ST_LINE_SEARCH;

    // ---- accept: primal, slack reset, duals with their own fraction-to-the-boundary step ----
    double a_d = 1.0;
    BO_UNROLL
    for (int i = 0; i < BO_MI; ++i)
      if (dz[i] < 0.0) a_d = fmin(a_d, -tau * z[i] / dz[i]);
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) x[i] = xt[i];
    BO_UNROLL
    for (int i = 0; i < BO_MI; ++i) {
      s[i] = fmax(st[i], cIt[i]);  // slack reset: lowers theta, never raises the barrier objective
      z[i] += a_d * dz[i];
      // keep z within a factor kappa_sigma of the central-path value mu/s (IPOPT eq. 16)
      z[i] = fmax(fmin(z[i], kappa_sigma * mu / s[i]), mu / (kappa_sigma * s[i]));
    }
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) y[j] += a_used * y_step[j];
    recalc_y = dw > 0.0;
  }
  S.it = it + 1;
  return -1;
}

// Convenience driver for one instance (used by the host-compiled test harness).
BO_DEVICE int bo_ipm_solve(bo_ipm_state& S, const bo_solver_params prm) {
  bo_ipm_init(S, prm);
  int status;
  do {
    status = bo_ipm_iterate(S, prm);
  } while (status < 0);
  return status;
}

#ifndef BO_HOST_SIM
// Persistent lanes with per-lane work fetching.  Iteration counts differ widely between instances
// (mean ~17, tail > 100 on the IK workload); a one-instance-per-thread launch would idle 31 lanes of
// a warp while its slowest instance finishes.  Instead every lane runs a small state machine:
// fetch an instance index from a global counter, initialise, then execute ONE interior-point
// iteration per trip round the loop; all lanes of a warp reconverge at the top of the loop, so the
// iteration body (the expensive, straight-line part) always runs warp-wide, and a lane whose
// instance has finished picks up the next one instead of waiting.
// Global layout: row-major [B][n] (instance-major), see b200optas.h.
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
      const int status = bo_ipm_iterate(S, prm);
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
