// bo_ipm_reg.cuh -- batched primal-dual interior-point / Newton-KKT solver, ONE PROBLEM INSTANCE PER
// THREAD (lane), all iterations inside one launch.
//
// Replaces, for a whole batch at once, what the reference does per call inside
// CasADiSolver._solve (optas/solver.py:386-398 -> casadi nlpsol("ipopt")): evaluate
// f, grad f, c, Jacobians and the Hessian of the Lagrangian, assemble the primal-dual KKT system,
// factor it (LDL' with inertia-correcting regularisation), filter line search, barrier update.
//
// Three tiers share this file (selected by bo_problem_create, see bo_api.cpp):
//   dense   (nx+n_eq <= 14)      generated straight-line tapes; KKT matrix packed, LDL' fully unrolled
//   sparse  (BO_SPARSE_LDL)      generated tapes; table-driven sparse LDL' from a host-side symbolic analysis
//   large   (BO_LARGE)           nothing generated: tapes interpreted, Jacobian products / KKT assembly from
//                                coordinate tables, factor values in a global [element][lane] scratch
//
// This header is included AFTER the generated prelude, which defines
//   BO_NX, BO_NP, BO_ME, BO_MI, BO_NNZ_JE, BO_NNZ_JI, BO_NNZ_H, BO_TPB
//   bo_tape_fc(x, p, f, cE, cI)
//   bo_tape_kkt(x, p, y, z, f, g, cE, cI, JE, JI, H)
//   bo_JEt_acc / bo_JIt_acc (out += J' v), bo_JE_mul / bo_JI_mul (out = J v),
//   bo_kkt_fill(H, JE, JI, sigma, K), bo_JEtJE_acc(JE, rho, K)
//
// Problem:  min f(x)  s.t.  cE(x) = 0,  cI(x) - s = 0,  s >= 0      (s: slacks)
// Lagrangian L = f - y'cE - z'cI,  z >= 0.  Barrier sub-problem parameter mu.
// Reduced Newton system solved each iteration (ds, dz eliminated):
//   [ H + JI' S JI + dw I    JE'   ] [ dx  ]   [ -(grad L) - JI' ((z*cI - mu)/s) ]
//   [ JE                    -dc I  ] [ -dy ] = [ -cE                            ],   S = diag(z/s)
#pragma once
#include "bo_common.cuh"

// Vector loops are unrolled (registers, compile-time indices) for small problems only; for larger
// ones unrolling 100+ element loops everywhere explodes code size and compile time for no benefit
// (the vectors live in local memory either way).
#if (BO_NX + BO_ME + BO_MI) > 64 && !defined(BO_HOST_SIM)
#undef BO_UNROLL
#define BO_UNROLL _Pragma("unroll 1")
#endif

#define BO_NK (BO_NX + BO_ME)
#ifdef BO_SPARSE_LDL
#define BO_KSZ BO_SPARSE_VALS /* D (BO_NK) followed by the structural non-zeros of L */
#else
#define BO_KSZ ((BO_NK * (BO_NK + 1)) / 2)
#endif
#define BO_KIDX(i, j) (((i) * ((i) + 1)) / 2 + (j)) /* packed lower triangle, i >= j */
#define BO_DIM(n) ((n) > 0 ? (n) : 1)
// A pivot of the x block counts as "not positive" below this fraction of its diagonal entry.  The interior-point kernel
// reads the inertia from it (a small pivot = a direction of negative / zero curvature: more dw).  The QP path knows its
// matrix is positive semidefinite; there a small ratio is only the usual ill-conditioning of Z/S near the solution, and
// regularising it away would spoil the Newton step: it accepts pivots down to rounding level.
#ifndef BO_PIVOT_RTOL
#ifdef BO_QP
#define BO_PIVOT_RTOL 1e-16
#else
#define BO_PIVOT_RTOL 1e-13
#endif
#endif

// ---- fast path: unpivoted LDL' with a fixed elimination order (x first, then y) ----
// Every lane executes the same instruction sequence (no data-dependent pivoting), the loops have
// compile-time bounds, and for small systems everything unrolls into straight-line FMAs.  It is only
// valid when the (1,1) block is positive definite; the caller makes it so WITHOUT changing the
// solution by adding rho*JE'JE to it and rho*JE'(rhs_y) to the right-hand side (adding rho*JE' times
// the second block row to the first).  By Debreu's lemma H + rho JE'JE is positive definite for large
// enough rho exactly when the reduced Hessian is, i.e. when the KKT inertia is correct.
// With a constraint-block perturbation dc the same trick needs dc' = dc / (1 - rho dc) and
// dy = dy'' / (1 - rho dc) (substitute dy'' = (1 - rho dc) dy to restore symmetry).
// On exit LD holds D on the diagonal and the unit-lower L below it.  The pivot signs ARE the inertia:
// a non-positive pivot in the x block means the reduced Hessian is not positive definite (or rho is
// too small, which is handled the same way: more dw), a non-negative one in the y block means JE is
// rank deficient (handled with dc) -- IPOPT's Algorithm IC, without a pivoting factorisation.
// Define BO_USE_BK (BO_FLAG_PIVOTED_LDL) to factor with Bunch-Kaufman partial pivoting instead.
#ifndef BO_DC_SCALE
#define BO_DC_SCALE 1e-8
#endif
#ifndef BO_STATIC_RHO
#define BO_STATIC_RHO 1.0e6
#endif
#ifdef BO_SPARSE_LDL
#define BO_LDL_SOLVE(LD, b) bo_ldl_sparse_solve(BO_LDP(S), BO_LDS(S), prm.ldl_tab, b)
#else
#define BO_LDL_SOLVE(LD, b) bo_ldl_static_solve(LD, b)
#endif
// small systems: unroll completely (indices become compile-time, the matrix lives in registers)
#if BO_NK <= 14
#define BO_LDL_UNROLL BO_UNROLL
#else
#define BO_LDL_UNROLL BO_NOUNROLL
#endif
BO_NOINLINE int bo_ldl_static(double* BO_RESTRICT A) {
  int bad = 0;  // 0 ok, 1 = non-positive pivot in the x block, 2 = non-negative pivot in the y block
  BO_LDL_UNROLL
  for (int j = 0; j < BO_NK; ++j) {
    double d = A[BO_KIDX(j, j)];
    const double scale = fmax(1.0, fabs(d));
    BO_LDL_UNROLL
    for (int k = 0; k < j; ++k) {
      const double l = A[BO_KIDX(j, k)];
      d -= l * l * A[BO_KIDX(k, k)];
    }
    if (j < BO_NX) {
      if (!(d > BO_PIVOT_RTOL * scale) && bad == 0) bad = 1;
    } else {
      if (!(d < -1e-13) && bad == 0) bad = 2;
    }
    A[BO_KIDX(j, j)] = d;
    const double dinv = 1.0 / d;
    BO_LDL_UNROLL
    for (int i = j + 1; i < BO_NK; ++i) {
      double v = A[BO_KIDX(i, j)];
      BO_LDL_UNROLL
      for (int k = 0; k < j; ++k) v -= A[BO_KIDX(i, k)] * A[BO_KIDX(j, k)] * A[BO_KIDX(k, k)];
      A[BO_KIDX(i, j)] = v * dinv;
    }
  }
  return bad;
}

BO_NOINLINE void bo_ldl_static_solve(const double* BO_RESTRICT A, double* BO_RESTRICT b) {
  BO_LDL_UNROLL
  for (int i = 1; i < BO_NK; ++i) {
    BO_LDL_UNROLL
    for (int k = 0; k < i; ++k) b[i] -= A[BO_KIDX(i, k)] * b[k];
  }
  BO_LDL_UNROLL
  for (int i = 0; i < BO_NK; ++i) b[i] /= A[BO_KIDX(i, i)];
  BO_LDL_UNROLL
  for (int i = BO_NK - 2; i >= 0; --i) {
    BO_LDL_UNROLL
    for (int k = i + 1; k < BO_NK; ++k) b[i] -= A[BO_KIDX(k, i)] * b[k];
  }
}

#ifdef BO_SPARSE_LDL
// ---- sparse variant of the unpivoted LDL': table-driven, identical control flow for every lane ----
// Tables (int32, built by bo_sparse.cpp at bo_problem_create, uploaded once):
//   [0] n  [1] nnzL  [2] off colptr[n+1]  [3] off rowidx[nnzL]  [4] off perm[n] (new -> old)
//   [5] off sign[n] (+1 variable, -1 constraint row)  [6] off factor program  [7] total length
// value array of one instance: vals[0..n) = D in elimination order, vals[n + e] = e-th non-zero of L.
// Factor program, column by column (left-looking):
//   n_k, then n_k records { pos L(j,k), pos D(k), cnt, cnt x { pos L(i,j), pos L(i,k) } }.
// `vals` is addressed with a stride so that the same code serves a thread-local array (stride 1) and
// the [element][lane] global scratch of the large tier (stride = number of lanes: coalesced).
#define BO_V(i) vals[(long long)(i) * stride]
BO_NOINLINE int bo_ldl_sparse(double* BO_RESTRICT vals, long long stride, const int32_t* BO_RESTRICT tab) {
  const int n = tab[0];
  const int32_t* colptr = tab + tab[2];
  const int32_t* sign = tab + tab[5];
  const int32_t* prog = tab + tab[6];
  int bad = 0, pc = 0;
  for (int j = 0; j < n; ++j) {
    double d = BO_V(j);
    const double scale = fmax(1.0, fabs(d));
    const int nk = prog[pc++];
    for (int kk = 0; kk < nk; ++kk) {
      const double ljk = BO_V(prog[pc]);
      const double w = ljk * BO_V(prog[pc + 1]);
      const int cnt = prog[pc + 2];
      pc += 3;
      d -= ljk * w;
      // targets (entries of column j) and sources (entries of column k) are disjoint and each target
      // appears once: four independent load/FMA/store chains in flight instead of one
      int c = 0;
      for (; c + 4 <= cnt; c += 4, pc += 8) {
        const int t0 = prog[pc], s0 = prog[pc + 1], t1 = prog[pc + 2], s1 = prog[pc + 3];
        const int t2 = prog[pc + 4], s2 = prog[pc + 5], t3 = prog[pc + 6], s3 = prog[pc + 7];
        const double a0 = BO_V(s0), a1 = BO_V(s1), a2 = BO_V(s2), a3 = BO_V(s3);
        const double v0 = BO_V(t0), v1 = BO_V(t1), v2 = BO_V(t2), v3 = BO_V(t3);
        BO_V(t0) = v0 - a0 * w;
        BO_V(t1) = v1 - a1 * w;
        BO_V(t2) = v2 - a2 * w;
        BO_V(t3) = v3 - a3 * w;
      }
      for (; c < cnt; ++c, pc += 2) BO_V(prog[pc]) -= BO_V(prog[pc + 1]) * w;
    }
    if (sign[j] > 0) {
      if (!(d > BO_PIVOT_RTOL * scale) && bad == 0) bad = 1;
    } else {
      if (!(d < -1e-13) && bad == 0) bad = 2;
    }
    BO_V(j) = d;
    const double dinv = 1.0 / d;
    for (int e = colptr[j]; e < colptr[j + 1]; ++e) BO_V(n + e) *= dinv;
  }
  return bad;
}

// b is indexed by ORIGINAL row (x then y); perm maps elimination position -> original row.
BO_NOINLINE void bo_ldl_sparse_solve(const double* BO_RESTRICT vals, long long stride, const int32_t* BO_RESTRICT tab,
                                     double* BO_RESTRICT b) {
  const int n = tab[0];
  const int32_t* colptr = tab + tab[2];
  const int32_t* rowidx = tab + tab[3];
  const int32_t* perm = tab + tab[4];
  for (int j = 0; j < n; ++j) {
    const double bj = b[perm[j]];
    for (int e = colptr[j]; e < colptr[j + 1]; ++e) b[perm[rowidx[e]]] -= BO_V(n + e) * bj;
  }
  for (int j = 0; j < n; ++j) b[perm[j]] /= BO_V(j);
  for (int j = n - 1; j >= 0; --j) {
    double acc = b[perm[j]];
    for (int e = colptr[j]; e < colptr[j + 1]; ++e) acc -= BO_V(n + e) * b[perm[rowidx[e]]];
    b[perm[j]] = acc;
  }
}
#undef BO_V
#endif

#ifdef BO_USE_BK
// Bunch-Kaufman LDL' (diagonal pivoting with 1x1 and 2x2 blocks; the unblocked LAPACK dsytf2
// algorithm, lower variant) of the symmetric indefinite KKT matrix, in place.  Pivoting is data
// dependent, so A lives in thread-local memory (L1-resident: 800 B for the 7-DoF IK problem).
// Outputs the pivot record ipiv (LAPACK convention, 1-based, negative for 2x2 blocks) and the
// inertia; returns false if a pivot block is numerically singular.
BO_NOINLINE bool bo_bk_factor(double* BO_RESTRICT A, int* BO_RESTRICT ipiv, int* n_neg_out) {
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
BO_NOINLINE void bo_bk_solve(const double* BO_RESTRICT A, const int* BO_RESTRICT ipiv, double* BO_RESTRICT b) {
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

// Factor A + diag(dw I, -dc I) in place (A = assembled KKT matrix, packed lower triangle).  Returns 0
// when the inertia is (BO_NX, BO_ME, 0), +1 when there are too many negative eigenvalues (reduced
// Hessian not positive definite), -1 when singular.
BO_DEVICE int bo_kkt_factor(double dw, double dc, double* BO_RESTRICT LD, int* BO_RESTRICT ipiv) {
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

#else
// stubs so that the (never taken) pivoted branches compile away
BO_DEVICE void bo_bk_solve(const double*, const int*, double*) {}
#endif

#ifdef BO_LARGE
#define BO_LDP(S) ((S).ldp)
#define BO_LDS(S) ((S).lds)
#else
#define BO_LDP(S) ((S).LD)
#define BO_LDS(S) 1LL
#endif

#ifdef BO_LARGE
// ===================== large tier: no problem-specific code, everything table-driven =====================
// For horizon problems (C4, C5: 10^5-instruction tapes, KKT systems of 10^3 rows) generating
// straight-line code is hopeless (megabytes of SASS, minutes of ptxas).  Instead the tapes are
// INTERPRETED: the instruction stream is read from global memory at a warp-uniform address
// (one broadcast transaction for 32 instances), operands live in a per-lane work array, and every lane
// takes the same branch of the opcode switch -- there is no divergence, because all instances share the
// tape.  Jacobian products and the KKT assembly run off coordinate tables the same way.
#include "bo_opcodes.h"

BO_NOINLINE void bo_tape_interp(const int32_t* BO_RESTRICT sec, const double* BO_RESTRICT dtab, double* BO_RESTRICT w,
                                const double* const* in, double* const* out) {
  const int n = sec[0];
  const double* consts = dtab + sec[1];
  // rows start 16-byte aligned (bo_sparse.cpp pads the section): one 128-bit load per instruction,
  // and the next row is fetched while the current one executes
  const bo_int4* rows = reinterpret_cast<const bo_int4*>(sec + 4);
  bo_int4 nxt = n > 0 ? rows[0] : bo_int4{0, 0, 0, 0};
  for (int i = 0; i < n; ++i) {
    const bo_int4 cur = nxt;
    if (i + 1 < n) nxt = rows[i + 1];
    const int op = cur.x & 0xFF, dst = cur.y, a = cur.z, b = cur.w;
    const int32_t ins0 = cur.x;
    double r;
    switch (op) {
      case BO_OP_INPUT: r = in[b][a]; break;
      case BO_OP_CONST: r = consts[a]; break;
      case BO_OP_OUTPUT: out[b][a] = w[dst]; continue;
      case BO_OP_ADD: r = w[a] + w[b]; break;
      case BO_OP_SUB: r = w[a] - w[b]; break;
      case BO_OP_MUL: r = w[a] * w[b]; break;
      case BO_OP_DIV: r = w[a] / w[b]; break;
      case BO_OP_NEG: r = -w[a]; break;
      case BO_OP_SQ: r = w[a] * w[a]; break;
      case BO_OP_SQRT: r = sqrt(w[a]); break;
      case BO_OP_SIN: r = sin(w[a]); break;
      case BO_OP_COS: r = cos(w[a]); break;
      case BO_OP_TAN: r = tan(w[a]); break;
      case BO_OP_ASIN: r = asin(w[a]); break;
      case BO_OP_ACOS: r = acos(w[a]); break;
      case BO_OP_ATAN: r = atan(w[a]); break;
      case BO_OP_ATAN2: r = atan2(w[a], w[b]); break;
      case BO_OP_FABS: r = fabs(w[a]); break;
      case BO_OP_FMIN: r = fmin(w[a], w[b]); break;
      case BO_OP_FMAX: r = fmax(w[a], w[b]); break;
      case BO_OP_EXP: r = exp(w[a]); break;
      case BO_OP_LOG: r = log(w[a]); break;
      case BO_OP_POW: r = pow(w[a], w[b]); break;
      case BO_OP_TANH: r = tanh(w[a]); break;
      case BO_OP_SINH: r = sinh(w[a]); break;
      case BO_OP_COSH: r = cosh(w[a]); break;
      case BO_OP_FLOOR: r = floor(w[a]); break;
      case BO_OP_CEIL: r = ceil(w[a]); break;
      case BO_OP_SIGN: r = bo_sign(w[a]); break;
      case BO_OP_NOT: r = (double)(w[a] == 0.0); break;
      case BO_OP_LT: r = (double)(w[a] < w[b]); break;
      case BO_OP_LE: r = (double)(w[a] <= w[b]); break;
      case BO_OP_EQ: r = (double)(w[a] == w[b]); break;
      case BO_OP_NE: r = (double)(w[a] != w[b]); break;
      case BO_OP_AND: r = (double)((w[a] != 0.0) && (w[b] != 0.0)); break;
      case BO_OP_OR: r = (double)((w[a] != 0.0) || (w[b] != 0.0)); break;
      case BO_OP_IF_ELSE: r = w[(unsigned)ins0 >> 8] != 0.0 ? w[a] : w[b]; break;
      default: r = BO_NAN;
    }
    w[dst] = r;
  }
}

BO_NOINLINE void bo_large_fc(const bo_solver_params& prm, const double* x, const double* p, double* f, double* cE, double* cI) {
  double w[BO_NWORK];
  const double* in[2] = {x, p};
  double* out[3] = {f, cE, cI};
  bo_tape_interp(prm.ldl_tab + prm.ldl_tab[16], prm.dtab, w, in, out);
}

BO_NOINLINE void bo_large_kkt(const bo_solver_params& prm, const double* x, const double* p, const double* y, const double* z,
                              double* f, double* g, double* cE, double* cI, double* JE, double* JI, double* H) {
  double w[BO_NWORK];
  const double* in[4] = {x, p, y, z};
  double* out[7] = {f, g, cE, cI, JE, JI, H};
  bo_tape_interp(prm.ldl_tab + prm.ldl_tab[17], prm.dtab, w, in, out);
}

// out[col] += J[k] * v[row]   /   out[row] = sum J[k] * x[col]
BO_NOINLINE void bo_coo_t_acc(const int32_t* BO_RESTRICT row, const int32_t* BO_RESTRICT col, int nnz, const double* J,
                              const double* v, double* out) {
  for (int k = 0; k < nnz; ++k) out[col[k]] += J[k] * v[row[k]];
}
BO_NOINLINE void bo_coo_mul(const int32_t* BO_RESTRICT row, const int32_t* BO_RESTRICT col, int nnz, int n_rows, const double* J,
                            const double* x, double* out) {
  for (int r = 0; r < n_rows; ++r) out[r] = 0.0;
  for (int k = 0; k < nnz; ++k) out[row[k]] += J[k] * x[col[k]];
}
BO_NOINLINE void bo_large_fill(const bo_solver_params& prm, const double* H, const double* JE, const double* JI,
                               const double* sigma, double* K, long long stride) {
  const int32_t* t = prm.ldl_tab;
  for (int i = 0; i < BO_KSZ; ++i) K[(long long)i * stride] = 0.0;
  const int32_t* hp = t + t[12];
  for (int k = 0; k < BO_NNZ_H; ++k) K[(long long)hp[k] * stride] += H[k];
  const int32_t* ip = t + t[14];
  const int ni = ip[0];
  ip += 1;
  for (int c = 0; c < ni; ++c, ip += 4) K[(long long)ip[0] * stride] += sigma[ip[3]] * JI[ip[1]] * JI[ip[2]];
  const int32_t* ep = t + t[13];
  for (int k = 0; k < BO_NNZ_JE; ++k) K[(long long)ep[k] * stride] += JE[k];
}
BO_NOINLINE void bo_large_jeje(const bo_solver_params& prm, const double* JE, double rho, double* K, long long stride) {
  const int32_t* ep = prm.ldl_tab + prm.ldl_tab[15];
  const int ne = ep[0];
  ep += 1;
  for (int c = 0; c < ne; ++c, ep += 3) K[(long long)ep[0] * stride] += rho * JE[ep[1]] * JE[ep[2]];
}
#define bo_eval_fc(x, p, f, cE, cI) bo_large_fc(prm, x, p, f, cE, cI)
#define bo_tape_kkt(x, p, y, z, f, g, cE, cI, JE, JI, H) bo_large_kkt(prm, x, p, y, z, f, g, cE, cI, JE, JI, H)
#define bo_JEt_acc(J, v, out) bo_coo_t_acc(prm.ldl_tab + prm.ldl_tab[8], prm.ldl_tab + prm.ldl_tab[9], BO_NNZ_JE, J, v, out)
#define bo_JIt_acc(J, v, out) bo_coo_t_acc(prm.ldl_tab + prm.ldl_tab[10], prm.ldl_tab + prm.ldl_tab[11], BO_NNZ_JI, J, v, out)
#define bo_JI_mul(J, x, out) bo_coo_mul(prm.ldl_tab + prm.ldl_tab[10], prm.ldl_tab + prm.ldl_tab[11], BO_NNZ_JI, BO_MI, J, x, out)
#define bo_JE_mul(J, x, out) bo_coo_mul(prm.ldl_tab + prm.ldl_tab[8], prm.ldl_tab + prm.ldl_tab[9], BO_NNZ_JE, BO_ME, J, x, out)
#define bo_kkt_fill(H, JE, JI, sigma, K) bo_large_fill(prm, H, JE, JI, sigma, K, BO_LDS(S))
#define bo_JEtJE_acc(JE, rho, K) bo_large_jeje(prm, JE, rho, K, BO_LDS(S))
#endif  // BO_LARGE

// Barrier objective and l1 constraint violation at (x, s) given the function values there.
BO_NOINLINE void bo_measures(double f, const double* cE, const double* cI, const double* s, double mu, double* phi,
                           double* theta) {
  double viol = 0.0, bar = 0.0;
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) viol += fabs(cE[j]);
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) {
    viol += fabs(cI[i] - s[i]);
    bar += log(s[i]);
  }
  *theta = viol;
  *phi = f - mu * bar;
}

#define BO_NFILTER 8
#ifndef BO_LS_MAX
#define BO_LS_MAX 16    /* backtracking halvings per direction (alpha >= 1.5e-5 alpha_max) */
#endif
#ifndef BO_HEAVY_MAX
#define BO_HEAVY_MAX 5  /* re-solves with dw = 1, 1e2, 1e4, 1e6 when no step is acceptable */
#endif
#define BO_IC_MAX 60    /* inertia-correction attempts per iteration */
#ifndef BO_REFINE_BELOW
#define BO_REFINE_BELOW 1e-4 /* refine the least-squares multipliers once the KKT error is below this */
#endif
#ifndef BO_INNER_ROUNDS
/* FACTOR/TRIAL repetitions per trip before the warp moves on to the next EVAL.  Pays when EVAL dominates
 * (sparse / large tiers: C4 1.7x faster); for the dense tier the phases cost about the same and waiting
 * for stragglers loses (measured on C2: 12.4 ms vs 8.9 ms), so there it is 1. */
#if defined(BO_SPARSE_LDL)
#define BO_INNER_ROUNDS 24
#else
#define BO_INNER_ROUNDS 1
#endif
#endif

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

#define BO_PH_EVAL 0    /* evaluate f, grad, c, J, H at x; test convergence; assemble K */
#define BO_PH_FACTOR 1  /* factor K + regularisation; on success compute the step */
#define BO_PH_TRIAL 2   /* evaluate one trial point; accept / correct / backtrack */

// Everything an instance carries between trips of the solver loop.  A GPU lane owns one of these
// and re-uses it for instance after instance (see bo_solve_kernel).
//
// The interior-point iteration is written as a *flattened state machine*: one call of
// bo_ipm_trip() runs at most ONE evaluation of the big tape, ONE factorisation and ONE trial-point
// evaluation, in that order, and every data-dependent retry loop of the algorithm (inertia
// correction, backtracking, second-order correction, convexified re-solve) is a transition that
// takes effect on the NEXT trip instead of an inner loop.  On the GPU all lanes of a warp execute
// trips in lock step, so the three heavy code blocks always run warp-wide: a lane that is
// backtracking costs its neighbours nothing, it simply uses the TRIAL block of the following trip
// while they are already in their next iteration.
struct bo_ipm_state {
  // instance data and iterate
  double p[BO_DIM(BO_NP)], x[BO_NX], s[BO_DIM(BO_MI)], y[BO_DIM(BO_ME)], z[BO_DIM(BO_MI)];
  double fth[BO_NFILTER], fph[BO_NFILTER];
  double f, mu, tau, dw_last, err0, theta_max, theta_min;
  int nf, it, n_acceptable, phase, trips;
  bool recalc_y, ls_mode, static_fac;  // static_fac: LD holds the unpivoted factorisation (rho-augmented)
  // evaluation at x (valid from PH_EVAL to the end of the iteration)
  double g[BO_NX], cE[BO_DIM(BO_ME)], cI[BO_DIM(BO_MI)], rd[BO_NX], sigma[BO_DIM(BO_MI)];
  double JE[BO_DIM(BO_NNZ_JE)], JI[BO_DIM(BO_NNZ_JI)], H[BO_DIM(BO_NNZ_H)];
#ifdef BO_LARGE
  double* ldp;        // this lane's slice of the global factor scratch ([element][lane])
  long long lds;
#else
  double LD[BO_KSZ];  // assembled KKT matrix, factored in place (re-assembled for every attempt)
#endif
#ifdef BO_USE_BK
  int ipiv[BO_NK];
#else
  int ipiv[1];
#endif
  double phi0, theta0, dw, dc, rho;
  int attempt, heavy;
  int n_singular;       // consecutive iterations whose unperturbed KKT matrix was singular (rank-deficient JE)
  bool jac_degenerate, first_singular;
  // step and line search
  double sol[BO_NK], dx[BO_NX], ds[BO_DIM(BO_MI)], y_step[BO_DIM(BO_ME)];
  double dx0[BO_NX], ds0[BO_DIM(BO_MI)], rE[BO_DIM(BO_ME)], rI[BO_DIM(BO_MI)];
  double a, a_trial, dphi, th_soc;
  int ls, soc;  // soc: 0 = plain trial, k > 0 = k-th second-order-corrected trial
  // feasibility restoration: iterations in the current phase (0: regular mode), phases so far, theta_r at x / at entry
  int resto, n_resto;
  double thr, thr0;
};

// The small tape is called from two places (start of an instance, every trial point): keep one copy.
#ifndef BO_LARGE
BO_NOINLINE void bo_eval_fc(const double* x, const double* p, double* f, double* cE, double* cI) { bo_tape_fc(x, p, f, cE, cI); }
#endif

#ifdef BO_QP
// quadratic cost, linear constraints: the dedicated iteration replaces everything below (same state, same linear algebra)
#include "bo_qp_reg.cuh"
#else
// Start an instance: S.p and S.x hold the parameters and the seed.
BO_DEVICE void bo_ipm_init(bo_ipm_state& S, const bo_solver_params prm) {
  S.mu = prm.mu_init;
  S.dw_last = 0.0;
  S.err0 = BO_INF;
  S.theta_max = BO_INF;
  S.theta_min = 0.0;
  S.nf = 0;
  S.it = 0;
  S.n_acceptable = 0;
  S.recalc_y = false;
  S.ls_mode = false;
  S.static_fac = false;
  S.n_singular = 0;
  S.jac_degenerate = false;
  S.phase = BO_PH_EVAL;
  S.trips = 0;
  S.resto = 0;
  S.n_resto = 0;
  bo_eval_fc(S.x, S.p, &S.f, S.cE, S.cI);
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) {
    S.s[i] = fmax(S.cI[i], 1e-2 * fmax(1.0, fabs(S.cI[i])));
    S.z[i] = S.mu / S.s[i];
  }
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) S.y[j] = 0.0;
}

// Restoration step data at x: weights / residuals of the violated inequality rows, no Hessian, no barrier, no multipliers.
// Returns theta_r.  (s, z are re-initialised when the restoration phase ends.)
BO_NOINLINE double bo_resto_prepare(bo_ipm_state& S) {
  double thr = 0.0;
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) {
    S.rE[j] = S.cE[j];
    thr += S.cE[j] * S.cE[j];
  }
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) {
    const double v = fmin(S.cI[i], 0.0);
    S.sigma[i] = v < 0.0 ? 1.0 : 0.0;
    S.rI[i] = v;
    S.z[i] = 0.0;
    S.s[i] = BO_INF;  // mu / s = 0, no fraction-to-the-boundary limit
    thr += v * v;
  }
  BO_UNROLL
  for (int i = 0; i < BO_NX; ++i) S.rd[i] = 0.0;
  BO_UNROLL
  for (int i = 0; i < BO_DIM(BO_NNZ_H); ++i) S.H[i] = 0.0;
  return sqrt(thr);
}

// Step for the constraint residuals (S.rE, S.rI) with the current factorisation: fills S.sol
// (dx, -dy), S.dx, S.ds and returns the fraction-to-the-boundary primal step length.
BO_NOINLINE double bo_ipm_step(bo_ipm_state& S, const bo_solver_params& prm) {
  double tvec[BO_DIM(BO_MI)];
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) tvec[i] = -(S.z[i] - S.mu / S.s[i] + S.sigma[i] * S.rI[i]);
  BO_UNROLL
  for (int i = 0; i < BO_NX; ++i) S.sol[i] = -S.rd[i];
  bo_JIt_acc(S.JI, tvec, S.sol);
  BO_UNROLL
  for (int j = 0; j < BO_ME; ++j) S.sol[BO_NX + j] = -S.rE[j];
  if (S.static_fac) {
    double t2[BO_DIM(BO_ME)];
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) t2[j] = -S.rho * S.rE[j];
    bo_JEt_acc(S.JE, t2, S.sol);  // first block row += rho * JE' * (second block rhs)
    BO_LDL_SOLVE(S.LD, S.sol);
    const double undo = 1.0 / (1.0 - S.rho * S.dc);
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) S.sol[BO_NX + j] *= undo;
  } else {
    bo_bk_solve(BO_LDP(S), S.ipiv, S.sol);
  }
  BO_UNROLL
  for (int i = 0; i < BO_NX; ++i) S.dx[i] = S.sol[i];
  bo_JI_mul(S.JI, S.dx, S.ds);
  double ap = 1.0;
  BO_NOUNROLL
  for (int i = 0; i < BO_MI; ++i) {
    S.ds[i] += S.rI[i];
    if (S.ds[i] < 0.0) ap = fmin(ap, -S.tau * S.s[i] / S.ds[i]);
  }
  return ap;
}

// One trip of the solver state machine = bo_trip_eval, bo_trip_factor, bo_trip_trial in that order.
// Each returns -1 to continue, or the final BO_ST_* status.  They are separate functions so that the
// GPU kernel can put a CTA barrier between them: the warps of a CTA then enter each (large,
// straight-line) block together and share its instruction fetch.
//
// Algorithm: primal-dual interior point with IPOPT's filter line search (Waechter & Biegler 2006,
// section 2.3, their default constants) on (theta = ||c||_1, phi = barrier objective), second-order
// correction of the first trial step (section 2.4), inertia-correcting regularisation (Algorithm
// IC).  IPOPT's restoration phase is replaced by re-solving the step with a heavily convexified
// Hessian (dw -> large turns it into the minimum-norm feasibility step).
BO_DEVICE int bo_trip_eval(bo_ipm_state& S, const bo_solver_params prm) {
  const double kappa_eps = 10.0, kappa_mu = 0.2, tau_min = 0.99, s_max = 100.0;
  const double mu_min = prm.tol * 0.1;
  // Trip budget.  Between an accepted step and the next evaluation S.f / S.err0 still describe the previous iterate:
  // an instance that runs out of budget there is evaluated once more so that what is reported belongs to the x returned.
  const bool over = ++S.trips > prm.max_trips;
  if (over && S.phase != BO_PH_EVAL) return BO_ST_MAX_ITER;

  // =========================== PH_EVAL ===========================
  if (S.phase == BO_PH_EVAL) {
    bo_tape_kkt(S.x, S.p, S.y, S.z, &S.f, S.g, S.cE, S.cI, S.JE, S.JI, S.H);
    if (S.resto > 0) {
      // restoration: next Levenberg-Marquardt step on the infeasibility from the fresh Jacobians
      if (!bo_isfinite(S.f)) return BO_ST_NUMERICAL;
      if (over || S.it >= prm.max_iter) return BO_ST_MAX_ITER;
      if (S.resto > BO_RESTO_MAX_IT) return BO_ST_LINE_SEARCH;
      S.thr = bo_resto_prepare(S);
      S.dw = BO_RESTO_ZETA;
      S.dc = BO_RESTO_DC;
      S.first_singular = false;
      S.attempt = 0;
      S.heavy = 0;
      S.ls_mode = false;
      S.phase = BO_PH_FACTOR;
    } else if (S.recalc_y && BO_ME > 0 && !over) {
      // The last step needed Hessian convexification (dw > 0): its Newton multipliers scale with dw
      // and feed back into the Hessian.  Replace y by the least-squares estimate
      //   [ I  JE' ; JE  -dc ] [ r ; y ] = [ grad f - JI' z ; 0 ]
      // (factor + solve in the FACTOR block below), then re-evaluate on the next trip.
      S.recalc_y = false;
      S.ls_mode = true;
      BO_UNROLL
      for (int i = 0; i < BO_DIM(BO_NNZ_H); ++i) S.H[i] = 0.0;
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) S.sigma[i] = 0.0;
      S.dw = 1.0;
      S.dc = 1e-10;
      S.phase = BO_PH_FACTOR;
    } else {
      // ---- residuals and the scaled optimality error (IPOPT's E_mu, Waechter & Biegler eq. 5) ----
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
      double e_dual = 0.0, e_prim = 0.0, e_comp0 = 0.0, sum_mult = 0.0, sum_z = 0.0;
      BO_UNROLL
      for (int i = 0; i < BO_NX; ++i) e_dual = fmax(e_dual, fabs(S.rd[i]));
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) {
        e_prim = fmax(e_prim, fabs(S.cE[j]));
        sum_mult += fabs(S.y[j]);
      }
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) {
        e_prim = fmax(e_prim, fabs(S.cI[i] - S.s[i]));
        e_comp0 = fmax(e_comp0, S.s[i] * S.z[i]);
        sum_z += fabs(S.z[i]);
      }
      sum_mult += sum_z;
      const double s_d = (BO_ME + BO_MI) > 0 ? fmax(s_max, sum_mult / (double)BO_DIM(BO_ME + BO_MI)) / s_max : 1.0;
      const double s_c = BO_MI > 0 ? fmax(s_max, sum_z / (double)BO_DIM(BO_MI)) / s_max : 1.0;
      S.err0 = fmax(fmax(e_dual / s_d, e_prim), e_comp0 / s_c);
#ifdef BO_HOST_TRACE
      printf("it %3d f %.6e err0 %.3e (dual %.3e prim %.3e comp %.3e) mu %.2e nf %d dw_last %.2e\n", S.it, S.f, S.err0,
             e_dual / s_d, e_prim, e_comp0 / s_c, S.mu, S.nf, S.dw_last);
#endif
      if (!bo_isfinite(S.err0) || !bo_isfinite(S.f)) return BO_ST_NUMERICAL;
      if (S.err0 <= prm.tol) return BO_ST_CONVERGED;
      S.n_acceptable = (S.err0 <= prm.acceptable_tol) ? S.n_acceptable + 1 : 0;
      if (S.n_acceptable >= 15) return BO_ST_ACCEPTABLE;
      if (S.it >= prm.max_iter || over) return BO_ST_MAX_ITER;

      // ---- barrier parameter update (monotone Fiacco-McCormick, IPOPT eq. 7); resets the filter ----
      if (BO_MI > 0) {
        for (int rep = 0; rep < 8; ++rep) {
          double e_comp = 0.0;
          BO_UNROLL
          for (int i = 0; i < BO_MI; ++i) e_comp = fmax(e_comp, fabs(S.s[i] * S.z[i] - S.mu));
          const double err_mu = fmax(fmax(e_dual / s_d, e_prim), e_comp / s_c);
          if (err_mu <= kappa_eps * S.mu && S.mu > mu_min) {
            S.mu = fmax(mu_min, fmin(kappa_mu * S.mu, S.mu * sqrt(S.mu)));  // mu^theta_mu, theta_mu = 1.5
            S.nf = 0;
          } else {
            break;
          }
        }
      }
      S.tau = fmax(tau_min, 1.0 - S.mu);
      BO_NOUNROLL
      for (int i = 0; i < BO_MI; ++i) S.sigma[i] = S.z[i] / S.s[i];
      bo_measures(S.f, S.cE, S.cI, S.s, S.mu, &S.phi0, &S.theta0);
      if (S.it == 0) {
        S.theta_max = 1e4 * fmax(1.0, S.theta0);
        S.theta_min = 1e-4 * fmax(1.0, S.theta0);
      }
      S.dw = 0.0;
      // IPOPT's degeneracy heuristic (PDPerturbationHandler): once the constraint Jacobian has been found rank
      // deficient in three consecutive iterations, the constraint block is perturbed from the first attempt on
      S.dc = S.jac_degenerate ? BO_DC_SCALE * sqrt(sqrt(S.mu)) : 0.0;
      S.first_singular = false;
      S.attempt = 0;
      S.heavy = 0;
      S.ls_mode = false;
      S.phase = BO_PH_FACTOR;
    }
  }

  return -1;
}

BO_DEVICE int bo_trip_factor(bo_ipm_state& S, const bo_solver_params prm) {
  // =========================== PH_FACTOR ===========================
  if (S.phase == BO_PH_FACTOR) {
    bo_kkt_fill(S.H, S.JE, S.JI, S.sigma, BO_LDP(S));
    int inertia;
#ifdef BO_USE_BK
    S.static_fac = false;
    inertia = bo_kkt_factor(S.dw, S.dc, BO_LDP(S), S.ipiv);
#else
    // unpivoted LDL' on the rho-augmented system (uniform control flow across the warp)
    S.static_fac = true;
    const double rho = S.ls_mode ? 0.0 : (S.resto > 0 ? 1.0 : BO_STATIC_RHO);
    S.rho = rho;
    if (!S.ls_mode) bo_JEtJE_acc(S.JE, rho, BO_LDP(S));
#ifdef BO_SPARSE_LDL
    {
      const int32_t* sign = prm.ldl_tab + prm.ldl_tab[5];
      const double dcp = S.dc / (1.0 - rho * S.dc);
      for (int j = 0; j < BO_NK; ++j) BO_LDP(S)[(long long)j * BO_LDS(S)] += sign[j] > 0 ? S.dw : -dcp;
    }
    const int bad = bo_ldl_sparse(BO_LDP(S), BO_LDS(S), prm.ldl_tab);
#else
    for (int i = 0; i < BO_NX; ++i) S.LD[BO_KIDX(i, i)] += S.dw;
    for (int i = BO_NX; i < BO_NK; ++i) S.LD[BO_KIDX(i, i)] -= S.dc / (1.0 - rho * S.dc);
    const int bad = bo_ldl_static(S.LD);
#endif
    inertia = bad == 0 ? 0 : (bad == 1 ? 1 : -1);
#endif
    if (S.ls_mode) {
      if (inertia == 0) {
        double nz[BO_DIM(BO_MI)];
        BO_UNROLL
        for (int i = 0; i < BO_MI; ++i) nz[i] = -S.z[i];
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) S.sol[i] = S.g[i];
        bo_JIt_acc(S.JI, nz, S.sol);
        BO_UNROLL
        for (int j = 0; j < BO_ME; ++j) S.sol[BO_NX + j] = 0.0;
        if (S.static_fac) BO_LDL_SOLVE(S.LD, S.sol);
        else bo_bk_solve(BO_LDP(S), S.ipiv, S.sol);
        bool fin = true;
        BO_UNROLL
        for (int j = 0; j < BO_ME; ++j) fin = fin && bo_isfinite(S.sol[BO_NX + j]);
        if (fin) {
          BO_UNROLL
          for (int j = 0; j < BO_ME; ++j) S.y[j] = S.sol[BO_NX + j];
          // One step of iterative refinement towards the UNregularised least-squares multipliers: the factorised
          // system carries -dc on the constraint block, which scales every component of y by s^2 / (s^2 + dc)
          // (s: singular value of JE) -- a relative error of dc / s^2 ~ 1e-8 that would otherwise sit in the dual
          // infeasibility as a floor just above tol (C4 stalled at 2.7e-8 and ended "acceptable").
          if (S.static_fac && S.err0 < BO_REFINE_BELOW) {  // only near convergence: elsewhere the accuracy buys nothing
            BO_UNROLL
            for (int i = 0; i < BO_NX; ++i) S.dx0[i] = S.sol[i];
            double ny[BO_DIM(BO_ME)];
            BO_UNROLL
            for (int j = 0; j < BO_ME; ++j) ny[j] = -S.y[j];
            BO_UNROLL
            for (int i = 0; i < BO_NX; ++i) S.sol[i] = S.g[i] - S.dw * S.dx0[i];
            bo_JIt_acc(S.JI, nz, S.sol);
            bo_JEt_acc(S.JE, ny, S.sol);
            bo_JE_mul(S.JE, S.dx0, S.rE);
            BO_UNROLL
            for (int j = 0; j < BO_ME; ++j) S.sol[BO_NX + j] = -S.rE[j];
            BO_LDL_SOLVE(S.LD, S.sol);
            BO_UNROLL
            for (int j = 0; j < BO_ME; ++j)
              if (bo_isfinite(S.sol[BO_NX + j])) S.y[j] += S.sol[BO_NX + j];
          }
        }
      }
      S.ls_mode = false;
      S.phase = BO_PH_EVAL;  // re-evaluate the Hessian with the new multipliers on the next trip
      return -1;
    }
#ifdef BO_HOST_TRACE
    if (inertia != 0) printf("     inertia %d at dw %.3e dc %.3e\n", inertia, S.dw, S.dc);
#endif
    if (inertia != 0) {
      // ---- inertia correction (IPOPT Algorithm IC); retried on the next trip ----
      if (inertia < 0 && BO_ME > 0 && S.dc == 0.0) {
        S.dc = BO_DC_SCALE * sqrt(sqrt(S.mu));  // IPOPT: 1e-8 mu^(1/4)  // singular: perturb the constraint block first
      if (S.attempt == 0 || BO_SINGULAR_ANY_ATTEMPT) S.first_singular = true;
      } else if (S.dw == 0.0) {
        S.dw = (S.dw_last == 0.0) ? 1e-4 : fmax(1e-20, S.dw_last / 3.0);
      } else {
        S.dw *= (S.dw_last == 0.0) ? 100.0 : 8.0;
      }
      if (++S.attempt > BO_IC_MAX || S.dw > 1e40) return S.err0 <= prm.acceptable_tol ? BO_ST_ACCEPTABLE : BO_ST_NUMERICAL;
      return -1;
    }
    if (S.dw > 0.0 && S.heavy == 0 && S.resto == 0) S.dw_last = S.dw;
    if (S.heavy == 0 && !S.jac_degenerate) {
      S.n_singular = S.first_singular ? S.n_singular + 1 : 0;
      if (S.n_singular >= 3) S.jac_degenerate = true;
    }
    if (S.resto == 0) {
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) S.rE[j] = S.cE[j];
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) S.rI[i] = S.cI[i] - S.s[i];
    }
    const double a_p = bo_ipm_step(S, prm);
    BO_UNROLL
    for (int j = 0; j < BO_ME; ++j) S.y_step[j] = -S.sol[BO_NX + j];
    double dphi = 0.0, dxn = 0.0;  // directional derivative of the barrier objective; step size
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) {
      dphi += S.g[i] * S.dx[i];
      dxn = fmax(dxn, fabs(S.dx[i]));
    }
    BO_NOUNROLL
    for (int i = 0; i < BO_MI; ++i) dphi -= S.mu * S.ds[i] / S.s[i];
    S.dphi = dphi;
    S.a = a_p;
    // step-length cap (cf. SNOPT's "major step limit"): Newton steps of several radians through
    // trigonometric kinematics are meaningless and wreck the multipliers
    if (prm.max_step > 0.0 && S.a * dxn > prm.max_step) S.a = prm.max_step / dxn;
    S.a_trial = S.a;
    S.ls = 0;
    S.soc = 0;
    S.phase = BO_PH_TRIAL;
  }

  return -1;
}

BO_DEVICE int bo_trip_trial(bo_ipm_state& S, const bo_solver_params prm) {
  const double kappa_sigma = 1e10, gamma_theta = 1e-5, gamma_phi = 1e-5, eta_phi = 1e-8;
  const double s_phi = 2.3, s_theta = 1.1, kappa_soc = 0.99;
  // =========================== PH_TRIAL ===========================
  if (S.phase == BO_PH_TRIAL) {
    double xt[BO_NX], st[BO_DIM(BO_MI)], cEt[BO_DIM(BO_ME)], cIt[BO_DIM(BO_MI)];
    double ft, phit, thetat;
    BO_UNROLL
    for (int i = 0; i < BO_NX; ++i) xt[i] = S.x[i] + S.a_trial * S.dx[i];
    BO_UNROLL
    for (int i = 0; i < BO_MI; ++i) st[i] = S.s[i] + S.a_trial * S.ds[i];
    bo_eval_fc(xt, S.p, &ft, cEt, cIt);
    if (S.resto > 0) {
      // restoration trial point: Armijo on theta_r
      double thr_t = 0.0;
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) thr_t += cEt[j] * cEt[j];
      BO_NOUNROLL
      for (int i = 0; i < BO_MI; ++i) {
        const double v = fmin(cIt[i], 0.0);
        thr_t += v * v;
      }
      thr_t = sqrt(thr_t);
#ifdef BO_HOST_TRACE
      printf("     resto %d ls %d a %.3e theta_r %.3e -> %.3e (entry %.3e)\n", S.resto, S.ls, S.a_trial, S.thr, thr_t, S.thr0);
#endif
      if (bo_isfinite(thr_t) && thr_t <= (1.0 - 1e-4 * S.a_trial) * S.thr) {
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) S.x[i] = xt[i];
        S.it += 1;
        if (thr_t <= fmax(BO_RESTO_KAPPA * S.thr0, 1e-10)) {
          // feasible enough: back to the regular iteration from here, multipliers and filter start afresh
          S.f = ft;
          BO_NOUNROLL
          for (int i = 0; i < BO_MI; ++i) {
            S.s[i] = fmax(cIt[i], 1e-2 * fmax(1.0, fabs(cIt[i])));
            S.z[i] = S.mu / S.s[i];
          }
          BO_UNROLL
          for (int j = 0; j < BO_ME; ++j) S.y[j] = 0.0;
          S.resto = 0;
          S.nf = 0;
          S.dw_last = 0.0;
          S.recalc_y = BO_ME > 0;  // least-squares equality multipliers at the new point
        } else {
          S.resto += 1;
        }
        S.phase = BO_PH_EVAL;
        return -1;
      }
      if (S.ls >= 2 && S.attempt < 4) {  // Levenberg-Marquardt: more damping, new direction
        S.attempt += 1;
        S.dw *= 100.0;
        S.phase = BO_PH_FACTOR;
        return -1;
      }
      S.a *= 0.5;
      S.a_trial = S.a;
      S.ls += 1;
      if (S.ls >= 24 || S.a < 1e-10) return BO_ST_LINE_SEARCH;  // stationary point of the infeasibility
      return -1;
    }
    bo_measures(ft, cEt, cIt, st, S.mu, &phit, &thetat);
    const bool finite = bo_isfinite(phit) && bo_isfinite(thetat);
    const bool ftype = S.dphi < 0.0 && S.theta0 <= S.theta_min &&
                       (S.theta0 <= 0.0 || log(S.a) + s_phi * log(-S.dphi) > s_theta * log(S.theta0));
    const double slack = 10.0 * 2.2e-16 * fabs(S.phi0);
    bool ok = false, armijo = false;
    if (finite && thetat <= S.theta_max) {
      bool in_filter = true;
      for (int j = 0; j < S.nf; ++j)
        if (!(thetat <= (1.0 - gamma_theta) * S.fth[j] || phit <= S.fph[j] - gamma_phi * S.fth[j])) in_filter = false;
      if (in_filter) {
        if (ftype) {
          armijo = phit - S.phi0 - slack <= eta_phi * S.a * S.dphi;
          ok = armijo;
        } else {
          ok = thetat <= (1.0 - gamma_theta) * S.theta0 || phit - slack <= S.phi0 - gamma_phi * S.theta0;
        }
      }
    }
#ifdef BO_HOST_TRACE
    printf("     heavy %d ls %d soc %d a %.3e ok %d ftype %d theta %.3e->%.3e phi %.8e->%.8e dphi %.3e dw %.2e\n", S.heavy, S.ls,
           S.soc, S.a_trial, (int)ok, (int)ftype, S.theta0, thetat, S.phi0, phit, S.dphi, S.dw);
#endif
    if (ok) {
      // ---- accept: primal, slack reset, duals with their own fraction-to-the-boundary step ----
      if (!(ftype && armijo)) {  // augment the filter (eq. 22)
        int slot = S.nf;
        if (S.nf < BO_NFILTER) {
          ++S.nf;
        } else {  // full: overwrite the entry with the largest theta
          slot = 0;
          for (int j = 1; j < BO_NFILTER; ++j)
            if (S.fth[j] > S.fth[slot]) slot = j;
        }
        S.fth[slot] = (1.0 - gamma_theta) * S.theta0;
        S.fph[slot] = S.phi0 - gamma_phi * S.theta0;
      }
      double a_d = 1.0, dz[BO_DIM(BO_MI)];
      BO_NOUNROLL
      for (int i = 0; i < BO_MI; ++i) {
        dz[i] = -S.z[i] + S.mu / S.s[i] - S.sigma[i] * S.ds[i];
        if (dz[i] < 0.0) a_d = fmin(a_d, -S.tau * S.z[i] / dz[i]);
      }
      BO_UNROLL
      for (int i = 0; i < BO_NX; ++i) S.x[i] = xt[i];
      BO_NOUNROLL
      for (int i = 0; i < BO_MI; ++i) {
        S.s[i] = fmax(st[i], cIt[i]);  // slack reset: lowers theta, never raises the barrier objective
        S.z[i] += a_d * dz[i];
        // keep z within a factor kappa_sigma of the central-path value mu/s (IPOPT eq. 16)
        S.z[i] = fmax(fmin(S.z[i], kappa_sigma * S.mu / S.s[i]), S.mu / (kappa_sigma * S.s[i]));
      }
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) S.y[j] += S.a * S.y_step[j];
      // regularised step (dw: nonconvex; dc: rank-deficient JE, whose multipliers are undetermined along
      // null(JE') and would otherwise drift by residual/dc): re-estimate y by least squares next trip
      // (The cooperative tier skips the re-estimate once the degeneracy heuristic keeps dc on for good, bo_ipm_cta.cuh: there a
      // factorisation costs more than the iterations it saves.  Here it does not: position + axis IK needs 6.0 iterations
      // with it and 7.8 without, 15.9 M against 8.8 M inst/s on B200.)
      #ifdef BO_RECALC_DC_ONLY  /* y does not enter the Hessian (linear equalities): only the rank-deficient case needs it */
    S.recalc_y = S.dc > 0.0;
#else
    S.recalc_y = S.dw > 0.0 || S.dc > 0.0;
#endif
      S.it += 1;
      S.phase = BO_PH_EVAL;
      return -1;
    }
    // ---- not acceptable ----
    bool try_soc = false;
    if (S.soc == 0) {
      // second-order correction (Waechter & Biegler section 2.4) for the first, full trial step only
      if (S.ls == 0 && finite && thetat >= S.theta0 && (BO_ME + BO_MI) > 0) {
        BO_UNROLL
        for (int i = 0; i < BO_NX; ++i) S.dx0[i] = S.dx[i];
        BO_UNROLL
        for (int i = 0; i < BO_MI; ++i) S.ds0[i] = S.ds[i];
        BO_UNROLL
        for (int j = 0; j < BO_ME; ++j) S.rE[j] = S.a * S.cE[j] + cEt[j];
        BO_UNROLL
        for (int i = 0; i < BO_MI; ++i) S.rI[i] = S.a * (S.cI[i] - S.s[i]) + (cIt[i] - st[i]);
        S.th_soc = thetat;
        try_soc = true;
      }
    } else if (S.soc < 4 && finite && thetat <= kappa_soc * S.th_soc) {
      BO_UNROLL
      for (int j = 0; j < BO_ME; ++j) S.rE[j] = S.a_trial * S.rE[j] + cEt[j];
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) S.rI[i] = S.a_trial * S.rI[i] + (cIt[i] - st[i]);
      S.th_soc = thetat;
      try_soc = true;
    }
    if (try_soc) {
      S.a_trial = bo_ipm_step(S, prm);  // corrected direction; tried on the next trip
      S.soc += 1;
      return -1;
    }
    if (S.soc > 0) {  // corrections did not help: back to the uncorrected direction
      BO_UNROLL
      for (int i = 0; i < BO_NX; ++i) S.dx[i] = S.dx0[i];
      BO_UNROLL
      for (int i = 0; i < BO_MI; ++i) S.ds[i] = S.ds0[i];
      S.soc = 0;
    }
    S.a *= 0.5;
    S.a_trial = S.a;
    S.ls += 1;
    if ((S.ls >= BO_LS_MAX || S.a < BO_ALPHA_MIN) && S.theta0 > 1e-7 * fmax(1.0, S.theta_min * 1e4) && S.n_resto < BO_RESTO_MAX_PHASES) {
      // no acceptable step at an infeasible point: feasibility restoration (the evaluation at x is still valid)
      S.n_resto += 1;
      S.resto = 1;
      S.thr0 = S.thr = bo_resto_prepare(S);
      S.dw = BO_RESTO_ZETA;
      S.dc = BO_RESTO_DC;
      S.attempt = 0;
      S.heavy = 0;
      S.phase = BO_PH_FACTOR;
      return -1;
    }
    if (S.ls >= BO_LS_MAX || S.a < BO_ALPHA_MIN) {
      // no acceptable step along this direction: convexify harder (dw large => minimum-norm
      // feasibility step); this stands in for IPOPT's restoration phase on these small problems
      if (++S.heavy >= BO_HEAVY_MAX) return S.err0 <= prm.acceptable_tol ? BO_ST_ACCEPTABLE : BO_ST_LINE_SEARCH;  // IPOPT: a failed step at an acceptable point ends "solved to acceptable level"
      S.dw = fmax(S.dw * 100.0, 1.0);
      S.phase = BO_PH_FACTOR;
    }
  }
  return -1;
}

BO_DEVICE int bo_ipm_trip(bo_ipm_state& S, const bo_solver_params prm) {
  int status = bo_trip_eval(S, prm);
  if (status < 0) status = bo_trip_factor(S, prm);
  if (status < 0) status = bo_trip_trial(S, prm);
  return status;
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
// Persistent lanes with per-lane work fetching.  Iteration counts differ widely between instances
// (mean ~17, tail > 100 on the IK workload); a one-instance-per-thread launch would idle 31 lanes of
// a warp while its slowest instance finishes.  Instead every lane runs a small state machine:
// fetch an instance index from a global counter, initialise, then execute one bo_ipm_trip() per
// trip round the loop; all lanes of a warp reconverge at the top of the loop, so the three heavy
// blocks of the trip (tape evaluation, factorisation, trial point) always run warp-wide, and a lane
// whose instance has finished picks up the next one instead of waiting.
// Global layout: row-major [B][n] (instance-major), see b200optas.h.
extern "C" __global__ void __launch_bounds__(BO_TPB)
bo_solve_kernel(long long B, const double* __restrict__ p_all, const double* __restrict__ x0_all,
                double* __restrict__ x_all, double* __restrict__ lam_all, double* __restrict__ f_all,
                int* __restrict__ status_all, int* __restrict__ iters_all, double* __restrict__ kkt_all,
                unsigned long long* __restrict__ work_counter, const bo_solver_params prm) {
  bo_ipm_state S;
#ifdef BO_LARGE
  S.ldp = prm.scratch + ((long long)blockIdx.x * BO_TPB + threadIdx.x);
  S.lds = prm.scratch_stride;
#endif
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
    // CTA-wide barrier per trip: keeps the warps of a CTA in the same block of this (large) code at
    // the same time, so instruction fetch is shared between them instead of thrashing the i-cache
    if (!__syncthreads_or(active ? 1 : 0)) break;
    int status = -1;
    if (active) status = bo_trip_eval(S, prm);
#ifndef BO_NO_PHASE_BARRIER
    __syncthreads();
#endif
    // In the sparse / large tiers FACTOR and TRIAL are much cheaper than EVAL.  Lanes whose
    // iteration needs several of them (inertia-correction retries, backtracking, second-order
    // corrections, convexified re-solves) repeat them HERE, warp by warp, until every lane of the warp
    // is ready for its next EVAL: the lanes of a warp then stay synchronised at iteration granularity
    // and the expensive block is not re-executed for the sake of a few stragglers.
    bool again = false;
    for (int round = 0; round < BO_INNER_ROUNDS; ++round) {
      if (round > 0 && again && ++S.trips > prm.max_trips) status = BO_ST_MAX_ITER;  // every extra round is a trip
      if (active && status < 0 && S.phase == BO_PH_FACTOR) status = bo_trip_factor(S, prm);
      if (active && status < 0 && S.phase == BO_PH_TRIAL) status = bo_trip_trial(S, prm);
      again = active && status < 0 && S.phase != BO_PH_EVAL;
      if (round + 1 >= BO_INNER_ROUNDS || !__any_sync(0xffffffffu, again)) break;
    }
    if (active) {
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
#endif  // BO_QP
