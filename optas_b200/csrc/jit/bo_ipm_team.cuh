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
//     and the filter.  ~2.9 KB per instance for C2 => 2 CTAs (64 instances, 8 warps) per SM.
//   * warp 0 is the MASTER of its 32 instances: it owns the scalar state machine in registers (phase, mu, filter
//     sizes ...), assembles and factors the KKT matrix in registers (unrolled, unpivoted LDL' on the rho-augmented
//     system -- see bo_ipm_reg.cuh for why that is valid) and takes every decision; the other warps only ever
//     execute slices, so no decision is computed twice with possibly different rounding.
//   * one trip of the loop = W1 (KKT tape slices) | M1 (residuals, convergence test, mu, factor, step) |
//     W2 (f / c slices at the trial point) | M2 (filter acceptance, SOC, backtracking), four CTA barriers.
//     As in bo_ipm_reg.cuh every data-dependent retry is a state transition that takes effect on the next trip.
//   * persistent CTAs; a team that finishes fetches the next instance (global counter) -- with TMA staging:
//     the CTA pulls TILES of 32 consecutive instances (p and x0 rows are contiguous byte ranges) into shared memory
//     with cp.async.bulk on an mbarrier, one tile ahead, and hands them out one by one.
//
// Generated prelude (bo_team.cpp) defines BO_NX, BO_NP, BO_ME, BO_MI, BO_NNZ_JE, BO_NNZ_JI, BO_NNZ_H, BO_G, BO_TPB,
// the tape slices bo_team_kkt(role, sm) / bo_team_fc(role, sm, at, rows) and the strided sparse helpers *_t.
#pragma once
#include "bo_common.cuh"

#define BO_NK (BO_NX + BO_ME)
#define BO_KSZ ((BO_NK * (BO_NK + 1)) / 2)
#define BO_KIDX(i, j) (((i) * ((i) + 1)) / 2 + (j)) /* packed lower triangle, i >= j */
#define BO_DIM(n) ((n) > 0 ? (n) : 1)

#ifndef BO_DC_SCALE
#define BO_DC_SCALE 1e-8
#endif
#ifndef BO_STATIC_RHO
#define BO_STATIC_RHO 1.0e6
#endif
#define BO_NFILTER 8
#ifndef BO_LS_MAX
#define BO_LS_MAX 16
#endif
#ifndef BO_HEAVY_MAX
#define BO_HEAVY_MAX 5
#endif
#define BO_IC_MAX 60
#ifndef BO_REFINE_BELOW
#define BO_REFINE_BELOW 1e-4
#endif

#define BO_PH_IDLE (-1)
#define BO_PH_EVAL 0
#define BO_PH_FACTOR 1
#define BO_PH_TRIAL 2
#define BO_PH_INIT 3

// ---- shared-memory layout: offsets in "elements" (one element = BO_LS doubles, one per lane) ----
#define BO_OFF_P 0
#define BO_OFF_X (BO_OFF_P + BO_NP)
#define BO_OFF_S (BO_OFF_X + BO_NX)
#define BO_OFF_Y (BO_OFF_S + BO_MI)
#define BO_OFF_Z (BO_OFF_Y + BO_ME)
#define BO_OFF_RS (BO_OFF_Z + BO_MI)      /* 1 / s */
#define BO_OFF_F0 (BO_OFF_RS + BO_MI)     /* f at x */
#define BO_OFF_G (BO_OFF_F0 + 1)
#define BO_OFF_CE (BO_OFF_G + BO_NX)
#define BO_OFF_CI (BO_OFF_CE + BO_ME)
#define BO_OFF_JE (BO_OFF_CI + BO_MI)
#define BO_OFF_JI (BO_OFF_JE + BO_NNZ_JE)
#define BO_OFF_H (BO_OFF_JI + BO_NNZ_JI)
#define BO_OFF_SIG (BO_OFF_H + BO_NNZ_H)
#define BO_OFF_RD (BO_OFF_SIG + BO_MI)
#define BO_OFF_LD (BO_OFF_RD + BO_NX)     /* packed factor: 1/D on the diagonal, unit-lower L below */
#define BO_OFF_DX (BO_OFF_LD + BO_KSZ)
#define BO_OFF_DS (BO_OFF_DX + BO_NX)
#define BO_OFF_YST (BO_OFF_DS + BO_MI)
#define BO_OFF_DX0 (BO_OFF_YST + BO_ME)
#define BO_OFF_DS0 (BO_OFF_DX0 + BO_NX)
#define BO_OFF_RE (BO_OFF_DS0 + BO_MI)
#define BO_OFF_RI (BO_OFF_RE + BO_ME)
#define BO_OFF_FT (BO_OFF_RI + BO_MI)     /* f at the trial point */
#define BO_OFF_CET (BO_OFF_FT + 1)
#define BO_OFF_CIT (BO_OFF_CET + BO_ME)
#define BO_OFF_PART (BO_OFF_CIT + BO_MI)  /* per role: sum log(s), sum |c| */
#define BO_OFF_AT (BO_OFF_PART + 2 * BO_G) /* trial step length, published by the master */
#define BO_OFF_FTH (BO_OFF_AT + 1)
#define BO_OFF_FPH (BO_OFF_FTH + BO_NFILTER)
#define BO_SM_ELEMS (BO_OFF_FPH + BO_NFILTER)

#ifdef BO_HOST_SIM
#define BO_LS 1
#else
#define BO_LS 32
#endif
#define SM(off, i) sm[((off) + (i)) * BO_LS]
#define SMP(off) (sm + (off) * BO_LS)

// Master-only scalar state of one instance (registers of warp 0).
struct bo_tm {
  double f, mu, tau, dw_last, err0, theta_max, theta_min, phi0, theta0, dw, dc, rho, a, a_trial, dphi, th_soc;
  int nf, it, n_acceptable, phase, trips, attempt, heavy, n_singular, ls, soc;
  bool recalc_y, ls_mode, jac_degenerate, first_singular;
  long long b;
};

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
    const double dinv = 1.0 / d;
    BO_UNROLL
    for (int i = j + 1; i < BO_NK; ++i) {
      double v = A[BO_KIDX(i, j)];
      BO_UNROLL
      for (int k = 0; k < j; ++k) v -= A[BO_KIDX(i, k)] * A[BO_KIDX(j, k)] * A[BO_KIDX(k, k)];
      A[BO_KIDX(i, j)] = v;  // unscaled C(i,j); scaled to L once column j is complete (below)
    }
    // rows > j of columns < j are still unscaled C; the products above used C(i,k) * C(j,k) / D(k) = L(i,k) C(j,k)
    A[BO_KIDX(j, j)] = dinv;
  }
  // scale C -> L
  BO_UNROLL
  for (int j = 0; j < BO_NK; ++j) {
    BO_UNROLL
    for (int i = j + 1; i < BO_NK; ++i) A[BO_KIDX(i, j)] *= A[BO_KIDX(j, j)];
  }
  return bad;
}

// Solve with the factor stored in shared memory (1/D on the diagonal): b is a register array.
BO_DEVICE void bo_tm_ldl_solve(const double* BO_RESTRICT sm, double* BO_RESTRICT b) {
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
}

// Step for the constraint residuals (RE, RI) with the current factorisation: writes DX, DS (and, when y_step is
// set, YST) and returns the fraction-to-the-boundary primal step length.
BO_NOINLINE double bo_tm_step(const bo_tm& M, double* BO_RESTRICT sm, const bool y_step) {
  double sol[BO_NK];
  BO_UNROLL
  for (int i = 0; i < BO_NX; ++i) sol[i] = -SM(BO_OFF_RD, i);
  // t = -(z - mu / s + sigma * rI), parked in DS (overwritten by the real ds below)
  BO_UNROLL
  for (int i = 0; i < BO_MI; ++i)
    SM(BO_OFF_DS, i) = -(SM(BO_OFF_Z, i) - M.mu * SM(BO_OFF_RS, i) + SM(BO_OFF_SIG, i) * SM(BO_OFF_RI, i));
  bo_JIt_acc_t(SMP(BO_OFF_JI), SMP(BO_OFF_DS), BO_LS, 1.0, sol, 1);
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) sol[BO_NX + j] = -SM(BO_OFF_RE, j);
  bo_JEt_acc_t(SMP(BO_OFF_JE), SMP(BO_OFF_RE), BO_LS, -M.rho, sol, 1);  // first block row += rho JE' (second block rhs)
  bo_tm_ldl_solve(sm, sol);
  const double undo = 1.0 / (1.0 - M.rho * M.dc);
  BO_UNROLL
  for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_DX, i) = sol[i];
  if (y_step) {
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_YST, j) = -sol[BO_NX + j] * undo;
  }
  bo_JI_mul_t(SMP(BO_OFF_JI), sol, 1, SMP(BO_OFF_DS), BO_LS);
  double worst = 0.0;  // max over rows of -ds / s
  BO_UNROLL
  for (int i = 0; i < BO_MI; ++i) {
    const double ds = SM(BO_OFF_DS, i) + SM(BO_OFF_RI, i);
    SM(BO_OFF_DS, i) = ds;
    worst = fmax(worst, -ds * SM(BO_OFF_RS, i));
  }
  return worst * 1.0 > M.tau ? M.tau / worst : 1.0;
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
  M.phase = BO_PH_INIT;
  BO_UNROLL
  for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_DX, i) = 0.0;  // the INIT evaluation reads x + 0 * dx
  SM(BO_OFF_AT, 0) = 0.0;
}

// ---- M1: EVAL part (results of the KKT slices are in shared memory) and FACTOR part ----
// Returns -1 to continue or the final status.
BO_NOINLINE int bo_tm_m1(bo_tm& M, double* BO_RESTRICT sm, const bo_solver_params& prm) {
  const double kappa_eps = 10.0, kappa_mu = 0.2, tau_min = 0.99, s_max = 100.0;
  const double mu_min = prm.tol * 0.1;
  const bool over = ++M.trips > prm.max_trips;
  if (over && M.phase != BO_PH_EVAL) return BO_ST_MAX_ITER;

  if (M.phase == BO_PH_EVAL) {
    M.f = SM(BO_OFF_F0, 0);
    if (M.recalc_y && BO_ME > 0 && !over) {
      // least-squares multiplier estimate after a regularised step (bo_ipm_reg.cuh): factor [I JE'; JE -dc]
      M.recalc_y = false;
      M.ls_mode = true;
      M.dw = 1.0;
      M.dc = 1e-10;
      M.phase = BO_PH_FACTOR;
    } else {
      double e_dual = 0.0, e_prim = 0.0, e_comp0 = 0.0, sum_mult = 0.0, sum_z = 0.0;
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
      for (int j = 0; j < BO_ME; ++j) {
        e_prim = fmax(e_prim, fabs(SM(BO_OFF_CE, j)));
        sum_mult += fabs(SM(BO_OFF_Y, j));
      }
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) {
        const double s = SM(BO_OFF_S, i), z = SM(BO_OFF_Z, i);
        e_prim = fmax(e_prim, fabs(SM(BO_OFF_CI, i) - s));
        e_comp0 = fmax(e_comp0, s * z);
        sum_z += fabs(z);
      }
      sum_mult += sum_z;
      const double s_d = (BO_ME + BO_MI) > 0 ? fmax(s_max, sum_mult / (double)BO_DIM(BO_ME + BO_MI)) / s_max : 1.0;
      const double s_c = BO_MI > 0 ? fmax(s_max, sum_z / (double)BO_DIM(BO_MI)) / s_max : 1.0;
      M.err0 = fmax(fmax(e_dual / s_d, e_prim), e_comp0 / s_c);
#ifdef BO_HOST_TRACE
      printf("it %3d f %.6e err0 %.3e (dual %.3e prim %.3e comp %.3e) mu %.2e nf %d dw_last %.2e\n", M.it, M.f, M.err0,
             e_dual / s_d, e_prim, e_comp0 / s_c, M.mu, M.nf, M.dw_last);
#endif
      if (!bo_isfinite(M.err0) || !bo_isfinite(M.f)) return BO_ST_NUMERICAL;
      if (M.err0 <= prm.tol) return BO_ST_CONVERGED;
      M.n_acceptable = (M.err0 <= prm.acceptable_tol) ? M.n_acceptable + 1 : 0;
      if (M.n_acceptable >= 15) return BO_ST_ACCEPTABLE;
      if (M.it >= prm.max_iter || over) return BO_ST_MAX_ITER;

      // barrier parameter update (monotone Fiacco-McCormick, Waechter & Biegler eq. 7); resets the filter
      if (BO_MI > 0) {
        for (int rep = 0; rep < 8; ++rep) {
          double e_comp = 0.0;
          BO_UNROLL
          for (int i = 0; i < BO_MI; ++i) e_comp = fmax(e_comp, fabs(SM(BO_OFF_S, i) * SM(BO_OFF_Z, i) - M.mu));
          const double err_mu = fmax(fmax(e_dual / s_d, e_prim), e_comp / s_c);
          if (err_mu <= kappa_eps * M.mu && M.mu > mu_min) {
            M.mu = fmax(mu_min, fmin(kappa_mu * M.mu, M.mu * sqrt(M.mu)));
            M.nf = 0;
          } else {
            break;
          }
        }
      }
      M.tau = fmax(tau_min, 1.0 - M.mu);
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) SM(BO_OFF_SIG, i) = SM(BO_OFF_Z, i) * SM(BO_OFF_RS, i);
      // barrier objective and l1 violation from the partial sums of the slices
      double lg = 0.0, th = 0.0;
      BO_UNROLL
      for (int r = 0; r < BO_G; ++r) {
        lg += SM(BO_OFF_PART, 2 * r);
        th += SM(BO_OFF_PART, 2 * r + 1);
      }
      M.theta0 = th;
      M.phi0 = M.f - M.mu * lg;
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
    const double rho = M.ls_mode ? 0.0 : BO_STATIC_RHO;
    M.rho = rho;
    int bad;
    {
      double K[BO_KSZ];
      bo_kkt_fill_t(SMP(BO_OFF_H), SMP(BO_OFF_JE), SMP(BO_OFF_JI), SMP(BO_OFF_SIG), M.ls_mode ? 0.0 : 1.0, rho, K);
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
      if (inertia == 0) {
        double sol[BO_NK];
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) sol[i] = SM(BO_OFF_G, i);
        bo_JIt_acc_t(SMP(BO_OFF_JI), SMP(BO_OFF_Z), BO_LS, -1.0, sol, 1);
        BO_UNROLL
        for (int j = 0; j < BO_ME; ++j) sol[BO_NX + j] = 0.0;
        bo_tm_ldl_solve(sm, sol);
        bool fin = true;
        BO_UNROLL
        for (int j = 0; j < BO_ME; ++j) fin = fin && bo_isfinite(sol[BO_NX + j]);
        if (fin) {
          BO_UNROLL
          for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_Y, j) = sol[BO_NX + j];
          if (M.err0 < BO_REFINE_BELOW) {
            // one step of iterative refinement towards the unregularised least-squares multipliers (bo_ipm_reg.cuh)
            BO_UNROLL
            for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_DX0, i) = sol[i];
            BO_UNROLL
            for (int i = 0; i < BO_NX; ++i) sol[i] = SM(BO_OFF_G, i) - M.dw * SM(BO_OFF_DX0, i);
            bo_JIt_acc_t(SMP(BO_OFF_JI), SMP(BO_OFF_Z), BO_LS, -1.0, sol, 1);
            bo_JEt_acc_t(SMP(BO_OFF_JE), SMP(BO_OFF_Y), BO_LS, -1.0, sol, 1);
            bo_JE_mul_t(SMP(BO_OFF_JE), SMP(BO_OFF_DX0), BO_LS, sol + BO_NX, 1);
            BO_UNROLL
            for (int j = 0; j < BO_ME; ++j) sol[BO_NX + j] = -sol[BO_NX + j];
            bo_tm_ldl_solve(sm, sol);
            BO_UNROLL
            for (int j = 0; j < BO_ME; ++j)
              if (bo_isfinite(sol[BO_NX + j])) SM(BO_OFF_Y, j) += sol[BO_NX + j];
          }
        }
      }
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
        if (M.attempt == 0) M.first_singular = true;
      } else if (M.dw == 0.0) {
        M.dw = (M.dw_last == 0.0) ? 1e-4 : fmax(1e-20, M.dw_last / 3.0);
      } else {
        M.dw *= (M.dw_last == 0.0) ? 100.0 : 8.0;
      }
      if (++M.attempt > BO_IC_MAX || M.dw > 1e40) return M.err0 <= prm.acceptable_tol ? BO_ST_ACCEPTABLE : BO_ST_NUMERICAL;
      return -1;
    }
    if (M.dw > 0.0 && M.heavy == 0) M.dw_last = M.dw;
    if (M.heavy == 0 && !M.jac_degenerate) {
      M.n_singular = M.first_singular ? M.n_singular + 1 : 0;
      if (M.n_singular >= 3) M.jac_degenerate = true;
    }
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_RE, j) = SM(BO_OFF_CE, j);
    BO_UNROLL
    for (int i = 0; i < BO_MI; ++i) SM(BO_OFF_RI, i) = SM(BO_OFF_CI, i) - SM(BO_OFF_S, i);
    const double a_p = bo_tm_step(M, sm, true);
    double dphi = 0.0, dxn = 0.0, dsr = 0.0;
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) {
      const double dx = SM(BO_OFF_DX, i);
      dphi += SM(BO_OFF_G, i) * dx;
      dxn = fmax(dxn, fabs(dx));
    }
    BO_UNROLL
    for (int i = 0; i < BO_MI; ++i) dsr += SM(BO_OFF_DS, i) * SM(BO_OFF_RS, i);
    M.dphi = dphi - M.mu * dsr;
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
BO_NOINLINE int bo_tm_m2(bo_tm& M, double* BO_RESTRICT sm, const bo_solver_params& prm) {
  const double kappa_sigma = 1e10, gamma_theta = 1e-5, gamma_phi = 1e-5, eta_phi = 1e-8;
  const double s_phi = 2.3, s_theta = 1.1, kappa_soc = 0.99;
  if (M.phase == BO_PH_INIT) {
    // start of an instance: slacks from c_I(x0) pushed into the interior, z on the central path, y = 0
    M.f = SM(BO_OFF_FT, 0);
    BO_UNROLL
    for (int i = 0; i < BO_MI; ++i) {
      const double ci = SM(BO_OFF_CIT, i);
      const double s = fmax(ci, 1e-2 * fmax(1.0, fabs(ci)));
      SM(BO_OFF_S, i) = s;
      SM(BO_OFF_Z, i) = M.mu / s;
    }
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_Y, j) = 0.0;
    M.phase = BO_PH_EVAL;
    return -1;
  }
  if (M.phase != BO_PH_TRIAL) return -1;
  const double at = M.a_trial;
  const double ft = SM(BO_OFF_FT, 0);
  double lg = 0.0, thetat = 0.0;
  BO_UNROLL
  for (int r = 0; r < BO_G; ++r) {
    lg += SM(BO_OFF_PART, 2 * r);
    thetat += SM(BO_OFF_PART, 2 * r + 1);
  }
  const double phit = ft - M.mu * lg;
  const bool finite = bo_isfinite(phit) && bo_isfinite(thetat);
  const bool ftype = M.dphi < 0.0 && M.theta0 <= M.theta_min &&
                     (M.theta0 <= 0.0 || log(M.a) + s_phi * log(-M.dphi) > s_theta * log(M.theta0));
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
    BO_UNROLL
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
    for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_X, i) = SM(BO_OFF_X, i) + at * SM(BO_OFF_DX, i);
    BO_UNROLL
    for (int i = 0; i < BO_MI; ++i) {
      const double st = SM(BO_OFF_S, i) + at * SM(BO_OFF_DS, i);
      const double s = fmax(st, SM(BO_OFF_CIT, i));  // slack reset: lowers theta, never raises the barrier objective
      SM(BO_OFF_S, i) = s;
      double z = SM(BO_OFF_Z, i) + a_d * SM(BO_OFF_RI, i);
      // keep z within a factor kappa_sigma of the central-path value mu/s (IPOPT eq. 16); the test is division-free
      const double sz = s * z;
      if (sz > kappa_sigma * M.mu) z = kappa_sigma * M.mu / s;
      else if (sz * kappa_sigma < M.mu) z = M.mu / (kappa_sigma * s);
      SM(BO_OFF_Z, i) = z;
    }
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
      for (int i = 0; i < BO_MI; ++i) SM(BO_OFF_DS0, i) = SM(BO_OFF_DS, i);
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_RE, j) = M.a * SM(BO_OFF_CE, j) + SM(BO_OFF_CET, j);
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) {
        const double s = SM(BO_OFF_S, i);
        SM(BO_OFF_RI, i) = M.a * (SM(BO_OFF_CI, i) - s) + (SM(BO_OFF_CIT, i) - (s + at * SM(BO_OFF_DS, i)));
      }
      M.th_soc = thetat;
      try_soc = true;
    }
  } else if (M.soc < 4 && finite && thetat <= kappa_soc * M.th_soc) {
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) SM(BO_OFF_RE, j) = at * SM(BO_OFF_RE, j) + SM(BO_OFF_CET, j);
    BO_UNROLL
    for (int i = 0; i < BO_MI; ++i)
      SM(BO_OFF_RI, i) = at * SM(BO_OFF_RI, i) + (SM(BO_OFF_CIT, i) - (SM(BO_OFF_S, i) + at * SM(BO_OFF_DS, i)));
    M.th_soc = thetat;
    try_soc = true;
  }
  if (try_soc) {
    M.a_trial = bo_tm_step(M, sm, false);  // corrected direction; tried on the next trip
    M.soc += 1;
    return -1;
  }
  if (M.soc > 0) {  // corrections did not help: back to the uncorrected direction
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) SM(BO_OFF_DX, i) = SM(BO_OFF_DX0, i);
    BO_UNROLL
    for (int i = 0; i < BO_MI; ++i) SM(BO_OFF_DS, i) = SM(BO_OFF_DS0, i);
    M.soc = 0;
  }
  M.a *= 0.5;
  M.a_trial = M.a;
  M.ls += 1;
  if (M.ls >= BO_LS_MAX || M.a < 1e-12) {
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
    if (M.phase == BO_PH_TRIAL || M.phase == BO_PH_INIT)
      for (int r = 0; r < BO_G; ++r) bo_team_fc(r, sm, M.phase == BO_PH_INIT ? 0.0 : M.a_trial, M.phase == BO_PH_TRIAL);
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
    if (role == 0) {
      if (M.phase == BO_PH_EVAL || M.phase == BO_PH_FACTOR || M.phase == BO_PH_TRIAL) {
        const int status = bo_tm_m1(M, sm, prm);
        if (status >= 0) finish(status);
      }
      fetch();  // a team that just finished starts its next instance in this trip's W2
      SM(BO_OFF_AT, 0) = M.a_trial;
      ctrl[lane] = M.phase;
    }
    __syncthreads();
    // ---- W2: f / c slices at the trial point (or at the seed of a fresh instance) ----
    {
      const int ph = ctrl[lane];
      if (ph == BO_PH_TRIAL || ph == BO_PH_INIT) bo_team_fc(role, sm, SM(BO_OFF_AT, 0), ph == BO_PH_TRIAL);
    }
    __syncthreads();
    // ---- M2 ----
    if (role == 0 && (M.phase == BO_PH_TRIAL || M.phase == BO_PH_INIT)) {
      const int status = bo_tm_m2(M, sm, prm);
      if (status >= 0) finish(status);
    }
  }
}
#endif
