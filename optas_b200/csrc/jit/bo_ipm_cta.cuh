// bo_ipm_cta.cuh -- batched primal-dual interior-point solver, ONE PROBLEM INSTANCE PER CTA
// (cooperative tier; the algorithm and its constants are those of bo_ipm_reg.cuh, which documents them).
//
// Replaces CasADiSolver._solve (optas/solver.py:386-398 -> casadi nlpsol("ipopt")) for horizon problems
// (MPC / trajectory optimisation: example/point_mass_mpc.py, figure_eight_plan.py, dual_arm.py), whose
// per-instance state does not fit a thread:
//   * the sparse LDL' factor of the KKT system lives in SHARED memory and is computed by all threads of
//     the CTA with a level-scheduled left-looking program in which every entry has exactly one owner
//     (a group of BO_LDL_G lanes) -- deterministic, no atomics; one barrier per elimination-tree level;
//   * the tapes are partitioned into independent sub-tapes (one per horizon stage), thread t interprets
//     sub-tape t; the instruction streams of the 32 lanes of a warp are interleaved so that a warp reads
//     512 contiguous bytes per step, and isomorphic stages run in lock step;
//   * parameter-only sub-expressions (e.g. the FK of the current configuration that anchors a path) are
//     evaluated once per instance;
//   * triangular solves run on warp 0 with warp-level barriers only; everything else is thread-parallel
//     gathers + block reductions with fixed summation order (bitwise reproducible for any batch split).
// All tables come from bo_coop.cpp; this file has no problem-specific code (only the BO_* sizes).
//
// The same text compiles for the host (BO_HOST_SIM: one "thread", barriers and reductions are no-ops);
// that build is the CPU test harness of the solver logic, never part of libb200optas.so.
#pragma once
#include "bo_common.cuh"
#include "bo_opcodes.h"
#define BO_OPX_NOP 5

#define BO_NK (BO_NX + BO_ME)
#define BO_DIM(n) ((n) > 0 ? (n) : 1)

#ifndef BO_DC_SCALE
#define BO_DC_SCALE 1e-8
#endif
#ifndef BO_STATIC_RHO
#define BO_STATIC_RHO 1.0e6
#endif
#define BO_NFILTER 8
// feasibility restoration: see bo_ipm_reg.cuh (same algorithm, same constants)
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
#define BO_RESTO_DC (1.0 - 1e-8)
#ifndef BO_ALPHA_MIN
#define BO_ALPHA_MIN 5e-7
#endif
#ifndef BO_LS_MAX
#define BO_LS_MAX 16
#endif
#ifndef BO_HEAVY_MAX
#define BO_HEAVY_MAX 5
#endif
#define BO_IC_MAX 60
#ifndef BO_REFINE_BELOW
#define BO_REFINE_BELOW 1e-4 /* refine the least-squares multipliers once the KKT error is below this */
#endif
#define BO_PH_EVAL 0
#define BO_PH_FACTOR 1
#define BO_PH_TRIAL 2

// header slots of the integer table (bo_coop.h)
#define CT_N 0
#define CT_NNZL 1
#define CT_PERM 2
#define CT_SIGN 3
#define CT_NLEV 4
#define CT_PROG_FAC 5
#define CT_PROG_FWD 6
#define CT_PROG_BWD 7
#define CT_ANT 8
#define CT_APOS 9
#define CT_APTR 10
#define CT_ATERM 11
#define CT_JE_RPTR 12
#define CT_JE_RENT 13
#define CT_JE_CPTR 14
#define CT_JE_CENT 15
#define CT_JI_RPTR 16
#define CT_JI_RENT 17
#define CT_JI_CPTR 18
#define CT_JI_CENT 19
#define CT_TAPE_FC 20
#define CT_TAPE_KKT 21
#define TS_NSUB 0
#define TS_CONST0 1
#define TS_NPE 2
#define TS_NPART 3
#define TS_PRE_N 4
#define TS_PRE_OFF 5
#define TS_LEN_OFF 6
#define TS_WBASE_OFF 7
#define TS_STREAM_OFF 8
#define TS_NRED 9
#define TS_RED_OFF 10
#define TS_PART_SEG 11
#define TS_PE_SEG 12
#define TS_GEN_NWARP 13
#define TS_GEN_WL_OFF 14

// ---------------------------------------------------------------------------------------------------
// thread / barrier / reduction primitives
// ---------------------------------------------------------------------------------------------------
#ifdef BO_HOST_SIM
#define BO_TID 0
#define BO_NT 1
static inline void bo_sync() {}
static inline void bo_syncwarp() {}
static inline double bo_shfl_xor(double v, int) { return v; }
#else
#define BO_TID ((int)threadIdx.x)
#define BO_NT BO_TPB
__device__ __forceinline__ void bo_sync() { __syncthreads(); }
__device__ __forceinline__ void bo_bar_warps(int n_warps) {  // barrier among warps 0..n_warps-1 (named barrier 1)
  if (n_warps > 1) asm volatile("bar.sync 1, %0;" ::"r"(n_warps * 32) : "memory");
  else __syncwarp();
}
// level barrier of the factor program that also tells every participating thread whether ANY of them saw a bad pivot in
// the level (one decision for all warps: the program is abandoned at the same level everywhere)
__device__ __forceinline__ bool bo_bar_warps_or(int n_warps, bool flag) {
  if (n_warps > 1) {
    unsigned out;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbar.red.or.pred p, 1, %2, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(out) : "r"((unsigned)flag), "r"(n_warps * 32) : "memory");
    return out != 0;
  }
  return __any_sync(0xffffffffu, flag);
}
__device__ __forceinline__ void bo_syncwarp() { __syncwarp(); }
__device__ __forceinline__ double bo_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
#endif
#define BO_PAR(i, n) for (int i = BO_TID; i < (n); i += BO_NT)

// Shared memory of a CTA (doubles): [ red | 4 (ints) | vals + 1 | bp + 1 | free ]  -- BO_SMEM_DOUBLES in all, sized by
// bo_coop.cpp.  On the device the arrays are addressed off the __shared__ symbol so that the compiler emits
// LDS/STS with folded offsets instead of generic loads through a pointer kept in a struct.
#define BO_SM_RED 0
#define BO_SM_INTS (5 * (BO_TPB / 32))
#define BO_SM_VALS (BO_SM_INTS + 4)
#define BO_SM_BP (BO_SM_VALS + BO_VALS + 1)
#define BO_SM_FREE (BO_SM_BP + BO_NK + 1)
#ifdef BO_HOST_SIM
#define BO_VALS_AT(C, i) ((C).vals[i])
#define BO_BP_AT(C, i) ((C).bp[i])
#define BO_VALS_P(C) ((C).vals)
#define BO_BP_P(C) ((C).bp)
#define BO_RED_P(C) ((C).red)
#define BO_WKKT_P(C) ((C).wkkt)
#define BO_WFC_P(C) ((C).wfc)
#else
extern __shared__ double bo_smem[];
// index the __shared__ array itself (never through a generic pointer: that costs an S2UR + address rebuild per access)
#define BO_VALS_AT(C, i) (bo_smem[BO_SM_VALS + (i)])
#define BO_BP_AT(C, i) (bo_smem[BO_SM_BP + (i)])
#define BO_VALS_P(C) (bo_smem + BO_SM_VALS)
#define BO_BP_P(C) (bo_smem + BO_SM_BP)
#define BO_RED_P(C) (bo_smem + BO_SM_RED)
#define BO_WKKT_P(C) (bo_smem + BO_SM_VALS)
#define BO_WFC_P(C) (bo_smem + BO_SM_FREE)
#endif

#define BO_RED_SUM 0
#define BO_RED_MAX 1
#define BO_RED_MIN 2
BO_DEVICE double bo_red_op(int op, double a, double b) { return op == BO_RED_SUM ? a + b : (op == BO_RED_MAX ? fmax(a, b) : fmin(a, b)); }

// Reduce K values at once over the CTA; every thread receives all results.  Fixed tree: lanes by
// xor-shuffle, then the warps in index order -- the same bits for any launch geometry with this BO_TPB.
template <int K>
BO_DEVICE void bo_reduce(double (&v)[K], const int (&op)[K], double* red) {
#ifndef BO_HOST_SIM
  for (int off = 16; off > 0; off >>= 1)
    for (int k = 0; k < K; ++k) v[k] = bo_red_op(op[k], v[k], bo_shfl_xor(v[k], off));
  if (BO_NT > 32) {
    const int warp = BO_TID >> 5, lane = BO_TID & 31;
    if (lane == 0)
      for (int k = 0; k < K; ++k) red[warp * K + k] = v[k];
    __syncthreads();
    for (int k = 0; k < K; ++k) {
      double r = red[k];
      for (int w = 1; w < BO_NT / 32; ++w) r = bo_red_op(op[k], r, red[w * K + k]);
      v[k] = r;
    }
    __syncthreads();
  }
#else
  (void)v; (void)op; (void)red;
#endif
}

// ---------------------------------------------------------------------------------------------------
// per-CTA workspace
// ---------------------------------------------------------------------------------------------------
// vectors of one instance, in a per-CTA slice of global scratch (L1/L2 resident: <= 150 KB per CTA)
#define BO_OFF_P 0
#define BO_OFF_X (BO_OFF_P + BO_NP)
#define BO_OFF_S (BO_OFF_X + BO_NX)
#define BO_OFF_Y (BO_OFF_S + BO_MI)
#define BO_OFF_Z (BO_OFF_Y + BO_ME)
#define BO_OFF_G (BO_OFF_Z + BO_MI)
#define BO_OFF_CE (BO_OFF_G + BO_NX)
#define BO_OFF_CI (BO_OFF_CE + BO_ME)
#define BO_OFF_RD (BO_OFF_CI + BO_MI)
#define BO_OFF_SIG (BO_OFF_RD + BO_NX)
#define BO_OFF_JE (BO_OFF_SIG + BO_MI)
#define BO_OFF_JI (BO_OFF_JE + BO_NNZ_JE)
#define BO_OFF_H (BO_OFF_JI + BO_NNZ_JI)
#define BO_OFF_SOL (BO_OFF_H + BO_NNZ_H)
#define BO_OFF_DX (BO_OFF_SOL + BO_NK)
#define BO_OFF_DS (BO_OFF_DX + BO_NX)
#define BO_OFF_YST (BO_OFF_DS + BO_MI)
#define BO_OFF_DX0 (BO_OFF_YST + BO_ME)
#define BO_OFF_DS0 (BO_OFF_DX0 + BO_NX)
#define BO_OFF_RE (BO_OFF_DS0 + BO_MI)
#define BO_OFF_RI (BO_OFF_RE + BO_ME)
#define BO_OFF_XT (BO_OFF_RI + BO_MI)
#define BO_OFF_ST (BO_OFF_XT + BO_NX)
#define BO_OFF_CET (BO_OFF_ST + BO_MI)
#define BO_OFF_CIT (BO_OFF_CET + BO_ME)
#define BO_OFF_TV (BO_OFF_CIT + BO_MI)
#define BO_OFF_T2 (BO_OFF_TV + BO_MI)
#define BO_OFF_PEF (BO_OFF_T2 + BO_ME)
#define BO_OFF_PEK (BO_OFF_PEF + BO_NPE_FC)
#define BO_OFF_PART (BO_OFF_PEK + BO_NPE_KKT)
#define BO_OFF_F (BO_OFF_PART + BO_NPART)
#define BO_SCRATCH_DOUBLES (BO_OFF_F + 8)

#if defined(BO_PROFILE) && !defined(BO_HOST_SIM)
#define BO_PROF_BEGIN() const long long bo_prof_t0 = clock64()
#define BO_PROF_END(k) do { if (BO_TID == 0) C.prof[k] += clock64() - bo_prof_t0; } while (0)
#define BO_PROF_COUNT(k) do { if (BO_TID == 0) C.prof[k] += 1; } while (0)
#elif defined(BO_HOST_SIM)
// host build: event counters only (0 kkt tape, 1 f/c tape, 2 assembly, 5 factorisations, 6 solves), read by tests/hostsim.py
static long long bo_host_prof[8];
#define BO_PROF_COUNT(k) do { bo_host_prof[k] += 1; } while (0)
#define BO_PROF_BEGIN() do { } while (0)
#define BO_PROF_END(k) do { if ((k) < 3) bo_host_prof[k] += 1; } while (0)
#else
#define BO_PROF_COUNT(k) do { } while (0)
#define BO_PROF_BEGIN() do { } while (0)
#define BO_PROF_END(k) do { } while (0)
#endif

struct bo_cta {
  long long* prof;  // shared: cycle counters per phase (BO_PROFILE builds)
  double* wkkt;     // shared: work arrays of the KKT tape, w[slot][thread] (aliases vals/bp: the factor is dead then)
  double* wfc;      // shared: work arrays of the f/c tape (free tail)
  double* W;        // global scratch slice of this CTA
  double* vals;     // shared: 1/D (elimination order) then the unscaled factor entries C = L D
  double* bp;       // shared: right-hand side / solution in elimination order
  double* red;      // shared: reduction scratch
  int* ibuf;        // shared: small integers
  const int32_t* tab;
  const double* dtab;
};

// scalar state of the iteration: every thread of the CTA holds an identical copy (all control flow is CTA-uniform)
struct bo_cta_state {
  double fth[BO_NFILTER], fph[BO_NFILTER];
  double f, mu, tau, dw_last, err0, theta_max, theta_min;
  int nf, it, n_acceptable, phase, trips;
  bool recalc_y, ls_mode;
  bool rhs_ready;       // sol holds the right-hand side of the coming solve (computed once per entry into the factor phase, not per inertia retry)
  double phi0, theta0, dw, dc, rho;
  int attempt, heavy;
  int n_singular;       // consecutive iterations whose unperturbed KKT matrix was singular (rank-deficient JE)
  bool jac_degenerate, first_singular;
  double a, a_trial, dphi, th_soc;
  int ls, soc;
  int resto, n_resto;  // feasibility restoration: iterations in the current phase (0: regular mode), phases so far
  double thr, thr0;    // theta_r at x / at entry
};

// ---------------------------------------------------------------------------------------------------
// tape interpreter: thread t runs sub-tape t
// ---------------------------------------------------------------------------------------------------
// w[slot * wstride]: the work array is either thread-local (wstride 1) or a column of the shared-memory array
// w[slot][thread] (wstride = a multiple of 32: conflict-free whatever slots the lanes of a warp touch).
// The rows of a lane are `stride` int4 apart; they are fetched BO_ROW_CHUNK at a time, one chunk ahead
// (independent loads in flight; the stream is sequential and L2 resident).
#define BO_ROW_CHUNK 8
BO_DEVICE void bo_interp_rows(const bo_int4* BO_RESTRICT rows, int n, int stride, const double* BO_RESTRICT consts,
                              double* BO_RESTRICT w, int wstride, const double* const* in, double* const* out) {
#define BO_W(i) w[(i) * wstride]
  bo_int4 nxt[BO_ROW_CHUNK];
  BO_UNROLL
  for (int u = 0; u < BO_ROW_CHUNK; ++u) nxt[u] = u < n ? rows[(long long)u * stride] : bo_int4{BO_OPX_NOP, 0, 0, 0};
  for (int base = 0; base < n; base += BO_ROW_CHUNK) {
    bo_int4 cur[BO_ROW_CHUNK];
    BO_UNROLL
    for (int u = 0; u < BO_ROW_CHUNK; ++u) cur[u] = nxt[u];
    BO_UNROLL
    for (int u = 0; u < BO_ROW_CHUNK; ++u) {
      const int i = base + BO_ROW_CHUNK + u;
      nxt[u] = i < n ? rows[(long long)i * stride] : bo_int4{BO_OPX_NOP, 0, 0, 0};
    }
    BO_UNROLL
    for (int u = 0; u < BO_ROW_CHUNK; ++u) {
      const int op = cur[u].x & 0xFF, dst = cur[u].y, a = cur[u].z, b = cur[u].w;
      double r;
      switch (op) {
        case BO_OPX_NOP: continue;
        case BO_OP_INPUT: r = in[b][a]; break;
        case BO_OP_CONST: r = consts[a]; break;
        case BO_OP_OUTPUT: out[b][a] = BO_W(dst); continue;
        case BO_OP_ADD: r = BO_W(a) + BO_W(b); break;
        case BO_OP_SUB: r = BO_W(a) - BO_W(b); break;
        case BO_OP_MUL: r = BO_W(a) * BO_W(b); break;
        case BO_OP_DIV: r = BO_W(a) / BO_W(b); break;
        case BO_OP_NEG: r = -BO_W(a); break;
        case BO_OP_SQ: r = BO_W(a) * BO_W(a); break;
        case BO_OP_SQRT: r = sqrt(BO_W(a)); break;
        case BO_OP_SIN: r = sin(BO_W(a)); break;
        case BO_OP_COS: r = cos(BO_W(a)); break;
        case BO_OP_TAN: r = tan(BO_W(a)); break;
        case BO_OP_ASIN: r = asin(BO_W(a)); break;
        case BO_OP_ACOS: r = acos(BO_W(a)); break;
        case BO_OP_ATAN: r = atan(BO_W(a)); break;
        case BO_OP_ATAN2: r = atan2(BO_W(a), BO_W(b)); break;
        case BO_OP_FABS: r = fabs(BO_W(a)); break;
        case BO_OP_FMIN: r = fmin(BO_W(a), BO_W(b)); break;
        case BO_OP_FMAX: r = fmax(BO_W(a), BO_W(b)); break;
        case BO_OP_EXP: r = exp(BO_W(a)); break;
        case BO_OP_LOG: r = log(BO_W(a)); break;
        case BO_OP_POW: r = pow(BO_W(a), BO_W(b)); break;
        case BO_OP_TANH: r = tanh(BO_W(a)); break;
        case BO_OP_SINH: r = sinh(BO_W(a)); break;
        case BO_OP_COSH: r = cosh(BO_W(a)); break;
        case BO_OP_FLOOR: r = floor(BO_W(a)); break;
        case BO_OP_CEIL: r = ceil(BO_W(a)); break;
        case BO_OP_SIGN: r = bo_sign(BO_W(a)); break;
        case BO_OP_NOT: r = (double)(BO_W(a) == 0.0); break;
        case BO_OP_LT: r = (double)(BO_W(a) < BO_W(b)); break;
        case BO_OP_LE: r = (double)(BO_W(a) <= BO_W(b)); break;
        case BO_OP_EQ: r = (double)(BO_W(a) == BO_W(b)); break;
        case BO_OP_NE: r = (double)(BO_W(a) != BO_W(b)); break;
        case BO_OP_AND: r = (double)((BO_W(a) != 0.0) && (BO_W(b) != 0.0)); break;
        case BO_OP_OR: r = (double)((BO_W(a) != 0.0) || (BO_W(b) != 0.0)); break;
        case BO_OP_IF_ELSE: r = BO_W((unsigned)cur[u].x >> 8) != 0.0 ? BO_W(a) : BO_W(b); break;
        default: r = BO_NAN;
      }
      BO_W(dst) = r;
    }
  }
#undef BO_W
}

// Run a partitioned tape: in/out are the segment pointers with room for the extra segments (in[pe_seg] =
// parameter-only values, out[part_seg] = partial sums).  ws/wstride: shared-memory work arrays (wstride 0: use a
// thread-local array of NW doubles).  Ends with a barrier.
template <int NW>
BO_DEVICE void bo_cta_tape_run(const bo_cta& C, int slot, const double** in, double** out, double* pe, double* part, double* ws,
                               int wstride) {
  const int32_t* sec = C.tab + C.tab[slot];
  const double* consts = C.dtab + sec[TS_CONST0];
  in[sec[TS_PE_SEG]] = pe;
  out[sec[TS_PART_SEG]] = part;
  const int nsub = sec[TS_NSUB];
  const int32_t* len = C.tab + sec[TS_LEN_OFF];
  const int32_t* wbase = C.tab + sec[TS_WBASE_OFF];
  const bo_int4* stream = reinterpret_cast<const bo_int4*>(C.tab + sec[TS_STREAM_OFF]);
#ifdef BO_GEN_TAPES
  // tapes compiled to straight-line code per component class (bo_coop.cpp): warp w runs its work list, lane l the l-th
  // member of each entry -- isomorphic components (the stages of the horizon) in lock step, values in registers
  (void)nsub; (void)len; (void)wbase; (void)stream; (void)ws; (void)wstride;
#ifdef BO_HOST_SIM
  for (int w = 0; w < sec[TS_GEN_NWARP]; ++w)
    for (int l = 0; l < 32; ++l) {
      if (slot == CT_TAPE_FC) bo_gen_fc(C.tab, sec, w, l, in, out, consts);
      else bo_gen_kkt(C.tab, sec, w, l, in, out, consts);
    }
#else
  if (slot == CT_TAPE_FC) bo_gen_fc(C.tab, sec, BO_TID >> 5, BO_TID & 31, in, out, consts);
  else bo_gen_kkt(C.tab, sec, BO_TID >> 5, BO_TID & 31, in, out, consts);
#endif
#else
#ifdef BO_HOST_SIM
  double wl[NW];
  for (int t = 0; t < nsub; ++t) bo_interp_rows(stream + wbase[t >> 5] + (t & 31), len[t], 32, consts, wl, 1, in, out);
  (void)ws; (void)wstride;
#else
  if (wstride > 0) {
    if (BO_TID < nsub) bo_interp_rows(stream + wbase[BO_TID >> 5] + (BO_TID & 31), len[BO_TID], 32, consts, ws + BO_TID, wstride, in, out);
  } else {
    double wl[NW];
    for (int t = BO_TID; t < nsub; t += BO_NT) bo_interp_rows(stream + wbase[t >> 5] + (t & 31), len[t], 32, consts, wl, 1, in, out);
  }
#endif
#endif
  bo_sync();
  const int nred = sec[TS_NRED];
  const int32_t* red = C.tab + sec[TS_RED_OFF];
  BO_PAR(r, nred) {
    const int32_t* e = red + 4 * r;
    double acc = part[e[2]];
    for (int k = 1; k < e[3]; ++k) acc += part[e[2] + k];
    out[e[0]][e[1]] = acc;
  }
  bo_sync();
}
BO_NOINLINE void bo_cta_tape_fc(const bo_cta& C, const double** in, double** out) {
  BO_PROF_BEGIN();
  bo_cta_tape_run<BO_NWORK_FC>(C, CT_TAPE_FC, in, out, C.W + BO_OFF_PEF, C.W + BO_OFF_PART, BO_WFC_P(C), BO_FC_WSTRIDE);
  BO_PROF_END(1);
}
BO_NOINLINE void bo_cta_tape_kkt(const bo_cta& C, const double** in, double** out) {
  BO_PROF_BEGIN();
  bo_cta_tape_run<BO_NWORK_KKT>(C, CT_TAPE_KKT, in, out, C.W + BO_OFF_PEK, C.W + BO_OFF_PART, BO_WKKT_P(C), BO_KKT_WSTRIDE);
  BO_PROF_END(0);
}

// once per instance: the parameter-only sub-expressions of both tapes
BO_NOINLINE void bo_cta_pre(const bo_cta& C) {
  double w[BO_NWORK_PRE];
  for (int which = 0; which < 2; ++which) {
    // one thread each (different warps when there are several), sequential rows
    const int owner = (which == 0 || BO_NT <= 32) ? 0 : 32;
    if (BO_TID != owner % BO_NT) continue;
    const int32_t* sec = C.tab + C.tab[which == 0 ? CT_TAPE_FC : CT_TAPE_KKT];
    const double* in[2] = {nullptr, C.W + BO_OFF_P};
    double* out[1] = {C.W + (which == 0 ? BO_OFF_PEF : BO_OFF_PEK)};
    bo_interp_rows(reinterpret_cast<const bo_int4*>(C.tab + sec[TS_PRE_OFF]), sec[TS_PRE_N], 1, C.dtab + sec[TS_CONST0], w, 1, in, out);
  }
  bo_sync();
}

BO_DEVICE double bo_cta_eval_fc(const bo_cta& C, const double* x, double* cE, double* cI) {
  const double* in[3] = {x, C.W + BO_OFF_P, nullptr};
  double* out[4] = {C.W + BO_OFF_F, cE, cI, nullptr};
  bo_cta_tape_fc(C, in, out);
  return C.W[BO_OFF_F];
}

BO_DEVICE double bo_cta_eval_kkt(const bo_cta& C) {
  double* W = C.W;
  const double* in[5] = {W + BO_OFF_X, W + BO_OFF_P, W + BO_OFF_Y, W + BO_OFF_Z, nullptr};
  double* out[8] = {W + BO_OFF_F, W + BO_OFF_G, W + BO_OFF_CE, W + BO_OFF_CI, W + BO_OFF_JE, W + BO_OFF_JI, W + BO_OFF_H, nullptr};
  bo_cta_tape_kkt(C, in, out);
  return W[BO_OFF_F];
}

// ---------------------------------------------------------------------------------------------------
// Jacobian products as gathers (one owner per output element)
// ---------------------------------------------------------------------------------------------------
// sum over the entries of list `r` of J[nz] * v[other]
BO_DEVICE double bo_gather(const int32_t* BO_RESTRICT ptr, const int32_t* BO_RESTRICT ent, int r, const double* BO_RESTRICT J,
                           const double* BO_RESTRICT v) {
  // J, v and the index lists live in the per-CTA global workspace / the tables (L2): four entries' loads are issued
  // together (the list is a chain of dependent L2 round trips otherwise); the sum keeps its order
  double acc = 0.0;
  int e = ptr[r];
  const int end = ptr[r + 1];
  for (; e + 4 <= end; e += 4) {
    int ia[4], ib[4];
    double a[4], b[4];
    BO_UNROLL
    for (int u = 0; u < 4; ++u) {
      ia[u] = ent[2 * (e + u)];
      ib[u] = ent[2 * (e + u) + 1];
    }
    BO_UNROLL
    for (int u = 0; u < 4; ++u) {
      a[u] = J[ia[u]];
      b[u] = v[ib[u]];
    }
    BO_UNROLL
    for (int u = 0; u < 4; ++u) acc += a[u] * b[u];
  }
  for (; e < end; ++e) acc += J[ent[2 * e]] * v[ent[2 * e + 1]];
  return acc;
}
#define BO_JE_T(c, v) bo_gather(C.tab + C.tab[CT_JE_CPTR], C.tab + C.tab[CT_JE_CENT], c, C.W + BO_OFF_JE, v)
#define BO_JI_T(c, v) bo_gather(C.tab + C.tab[CT_JI_CPTR], C.tab + C.tab[CT_JI_CENT], c, C.W + BO_OFF_JI, v)
#define BO_JI_ROW(r, v) bo_gather(C.tab + C.tab[CT_JI_RPTR], C.tab + C.tab[CT_JI_RENT], r, C.W + BO_OFF_JI, v)

// ---------------------------------------------------------------------------------------------------
// KKT assembly + level-scheduled sparse LDL' in shared memory
// ---------------------------------------------------------------------------------------------------
// vals <- K + diag(dw I, -dcp I): every position has one owner thread.  Ends with a barrier.
BO_NOINLINE void bo_cta_assemble(const bo_cta& C, double rho, double dw, double dcp) {
  BO_PROF_BEGIN();
  double* vals = BO_VALS_P(C);
  BO_PAR(i, BO_VALS) vals[i] = 0.0;
  bo_sync();
  const int32_t* t = C.tab;
  const int nt = t[CT_ANT];
  const int32_t* apos = t + t[CT_APOS];
  const int32_t* aptr = t + t[CT_APTR];
  const bo_int4* term = reinterpret_cast<const bo_int4*>(t + t[CT_ATERM]);
  const int32_t* sign = t + t[CT_SIGN];
  const double *H = C.W + BO_OFF_H, *JE = C.W + BO_OFF_JE, *JI = C.W + BO_OFF_JI, *sigma = C.W + BO_OFF_SIG;
#ifdef BO_ASM_INTERLEAVE  /* short term lists (C5: 1.7 per position): bo_coop.cpp decides */
  // Everything read here (tables, H, JE, JI, sigma) is an L2 round trip, and a position's terms are a chain of them (its
  // list bounds -> a table row -> that row's operands).  Each thread therefore walks FOUR positions at once, step by step,
  // so that four independent chains are in flight; every position still adds its terms in list order.  A term has the form
  // (c a) b with c = sigma | rho | 1 and b = 1 for the plain entries.
  for (int k0 = BO_TID; k0 < nt; k0 += 4 * BO_NT) {
    int pos[4], e[4], end[4];
    double acc[4];
    int longest = 0;
    BO_UNROLL
    for (int u = 0; u < 4; ++u) {
      const int k = k0 + u * BO_NT;
      const bool valid = k < nt;
      pos[u] = valid ? apos[k] : -1;
      e[u] = valid ? aptr[k] : 0;
      end[u] = valid ? aptr[k + 1] : 0;
      acc[u] = 0.0;
      longest = end[u] - e[u] > longest ? end[u] - e[u] : longest;
    }
    for (int step = 0; step < longest; ++step) {
      bo_int4 q[4];
      double a[4], b[4], c[4];
      BO_UNROLL
      for (int u = 0; u < 4; ++u) q[u] = e[u] < end[u] ? term[e[u]] : bo_int4{-1, 0, 0, 0};
      BO_UNROLL
      for (int u = 0; u < 4; ++u) {
        const double* src = q[u].x == 0 ? H : (q[u].x == 2 ? JI : JE);
        a[u] = q[u].x >= 0 ? src[q[u].y] : 0.0;
        b[u] = q[u].x >= 2 ? src[q[u].z] : 1.0;
        c[u] = q[u].x == 2 ? sigma[q[u].w] : (q[u].x == 3 ? rho : 1.0);
      }
      BO_UNROLL
      for (int u = 0; u < 4; ++u) {
        if (q[u].x >= 0) acc[u] += c[u] * a[u] * b[u];
        e[u] += 1;
      }
    }
    BO_UNROLL
    for (int u = 0; u < 4; ++u)
      if (pos[u] >= 0) {
        if (pos[u] < BO_NK) acc[u] += sign[pos[u]] > 0 ? dw : -dcp;
        vals[pos[u]] = acc[u];
      }
  }
#else  /* long term lists (C4: the J_I' Sigma J_I products; C3): four terms of ONE position at a time */
  BO_PAR(k, nt) {
    const int pos = apos[k];
    double acc = 0.0;
    int e = aptr[k];
    const int end = aptr[k + 1];
    // four terms at a time: their table rows, then their operands, are loaded together (everything here is an L2 round
    // trip); every term has the form (c a) b with c = sigma | rho | 1 and b = 1 for the plain entries, summed in order
    for (; e + 4 <= end; e += 4) {
      bo_int4 q[4];
      double a[4], b[4], c[4];
      BO_UNROLL
      for (int u = 0; u < 4; ++u) q[u] = term[e + u];
      BO_UNROLL
      for (int u = 0; u < 4; ++u) {
        const double* src = q[u].x == 0 ? H : (q[u].x == 2 ? JI : JE);
        a[u] = src[q[u].y];
        b[u] = q[u].x >= 2 ? src[q[u].z] : 1.0;
        c[u] = q[u].x == 2 ? sigma[q[u].w] : (q[u].x == 3 ? rho : 1.0);
      }
      BO_UNROLL
      for (int u = 0; u < 4; ++u) acc += c[u] * a[u] * b[u];
    }
    for (; e < end; ++e) {
      const bo_int4 q = term[e];
      if (q.x == 0) acc += H[q.y];
      else if (q.x == 1) acc += JE[q.y];
      else if (q.x == 2) acc += sigma[q.w] * JI[q.y] * JI[q.z];
      else acc += rho * JE[q.y] * JE[q.z];
    }
    if (pos < BO_NK) acc += sign[pos] > 0 ? dw : -dcp;
    vals[pos] = acc;
  }
#endif
  if (BO_TID == 0) {  // padding operands of the lane programs
    vals[BO_VALS] = 0.0;
    BO_BP_P(C)[BO_NK] = 0.0;
  }
  bo_sync();
  BO_PROF_END(2);
}

// Lane programs (bo_coop.cpp): warp-wide pre-scheduled streams of fixed-size PACKETS, 8 bytes per lane and word, word i
// of lane l at [i][l].  A packet = one header word + PK operand words; a ROUND (one target per group of G lanes) is one
// or more packets:
//   header:  x = flags (FIRST: the round starts here | LAST: it ends here | LEVEL_END | EMPTY) | log2(G) << 4,  y = tgt | POSITIVE << 15
//            (tgt = 0x7FFF: this lane finalises nothing; the flags are the same in all lanes of the warp)
//   operand: x = a | b << 16,          y = c                 (padding operands address the always-zero cells)
// MODE 0 factor:   acc += vals[a] vals[b] vals[c];   finish: d = vals[tgt] - acc, diagonal: vals[tgt] = 1/d (+ pivot test)
// MODE 1 forward:  acc += vals[a] bp[b];             finish: bp[tgt] = (bp[tgt] - acc) vals[tgt]
// MODE 2 backward: acc += vals[a] bp[b];             finish: bp[tgt] -= acc vals[tgt]
// After the LAST packet of a round the G lanes of a group add up their accumulators (xor-shuffles) and lane 0 of the group
// finalises; after a LEVEL_END packet the participating warps synchronise.
// Why packets: the first version walked a flat word stream with a per-word "header or operand?" branch, so every operand
// cost a full dependent chain (word decode -> 3 shared-memory loads -> DMUL -> DFMA, ~65 cycles) and a level ~2300 cycles
// (C4: 394 levels, 898 k cycles per factorisation, 0.5 % of the FP64 peak).  With a fixed packet shape the loop body is
// branch-free: the 3 PK operand loads of a packet are issued back to back, the products are summed as a tree, the words
// themselves are prefetched BO_LP_DEPTH packets ahead (they do not depend on the data; the stream is L2 resident).
#ifndef BO_FAC_PK
#define BO_FAC_PK 8
#endif
#ifndef BO_FWD_PK
#define BO_FWD_PK 2
#endif
#ifndef BO_BWD_PK
#define BO_BWD_PK 2
#endif
#ifndef BO_LP_DEPTH
#define BO_LP_DEPTH 4
#endif
#define BO_LP_PAD_PACKETS 8 /* padding packets at the end of every warp stream (bo_coop.cpp emits them): >= BO_LP_DEPTH */
#define BO_PKT_FIRST 1u
#define BO_PKT_LAST 2u
#define BO_PKT_LEVEL_END 4u
#define BO_PKT_EMPTY 8u /* no operands in this packet (a warp with nothing to do in a level): only the barrier */

#ifdef BO_HOST_SIM
static inline double bo_cta_rcp(double d) { return 1.0 / d; }
#else
// reciprocal without the IEEE slow path (hardware seed + two Newton steps, within 2 ulp for normal arguments): one per pivot
__device__ __forceinline__ double bo_cta_rcp(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  if (r != 0.0 && fabs(r) < BO_INF) {
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
  }
  return r;
}
#endif

// sum of the PK products of one packet (two interleaved partial sums: the same order on the device and in the host build)
template <int MODE, int PK>
BO_DEVICE double bo_pkt_sum(const bo_cta& C, const bo_int2* w) {
  double e = 0.0, o = 0.0;
  BO_UNROLL
  for (int k = 0; k < PK; ++k) {
    const unsigned x = (unsigned)w[k].x, y = (unsigned)w[k].y;
    const double p = MODE == 0 ? BO_VALS_AT(C, x & 0xFFFFu) * BO_VALS_AT(C, x >> 16) * BO_VALS_AT(C, y)
                               : BO_VALS_AT(C, x & 0xFFFFu) * BO_BP_AT(C, x >> 16);
    if (k & 1) o += p;
    else e += p;
  }
  return e + o;
}

// old values of the target, loaded before the group sum so that their latency hides behind the shuffles
template <int MODE>
BO_DEVICE void bo_lane_target_load(const bo_cta& C, int tgt, double* v0, double* b0) {
  if (MODE == 0) {
    *v0 = BO_VALS_AT(C, tgt != 0x7FFF ? tgt : BO_VALS);  // targets beyond BO_VALS are right-hand-side entries (bp follows vals)
    *b0 = 0.0;
  } else {
    *v0 = BO_VALS_AT(C, tgt < BO_NK ? tgt : BO_VALS);
    *b0 = BO_BP_AT(C, tgt < BO_NK ? tgt : BO_NK);
  }
}

template <int MODE>
BO_DEVICE bool bo_lane_finish(const bo_cta& C, int tgt, bool positive, double acc, double v0, double b0) {
  bool bad = false;
  if (MODE == 0) {
    const double d = v0 - acc;
    if (tgt < BO_NK) {
      const double scale = fmax(1.0, fabs(v0));
      bad = positive ? !(d > 1e-13 * scale) : !(d < -1e-13);
#ifdef BO_HOST_SIM
      if (bad && tgt < C.ibuf[0]) C.ibuf[0] = tgt;
#else
      if (bad) atomicMin(reinterpret_cast<int*>(&bo_smem[BO_SM_INTS]), tgt);
#endif
      BO_VALS_AT(C, tgt) = bo_cta_rcp(d);
    } else {
      BO_VALS_AT(C, tgt) = d;
    }
  } else if (MODE == 1) {
    BO_BP_AT(C, tgt) = (b0 - acc) * v0;
  } else {
    BO_BP_AT(C, tgt) = b0 - acc * v0;
  }
  return bad;
}

template <int MODE, int G, int PK>
BO_DEVICE void bo_lane_program(const bo_cta& C, int slot) {
  const int32_t* h = C.tab + C.tab[slot];
  const int W = h[0];
#ifdef BO_HOST_SIM
  // faithful emulation: warps advance level by level, lanes in lock step, the same xor-tree for the group sums
  int pc[32] = {0};
  static double acc[32][32];
  bool more = true;
  while (more) {
    more = false;
    for (int w = 0; w < W; ++w) {
      const int32_t* s = C.tab + h[2] + 64LL * h[4 + 2 * w];
      const int n = h[5 + 2 * w] / (PK + 1);
      while (pc[w] < n) {
        const int32_t* pkt = s + 64LL * (PK + 1) * pc[w];
        const unsigned hx = (unsigned)pkt[0];
        for (int l = 0; l < 32; ++l) {
          bo_int2 ops[PK];
          for (int k = 0; k < PK; ++k) ops[k] = bo_int2{pkt[64 * (k + 1) + 2 * l], pkt[64 * (k + 1) + 2 * l + 1]};
          const double sum = bo_pkt_sum<MODE, PK>(C, ops);
          acc[w][l] = (hx & BO_PKT_FIRST) ? sum : acc[w][l] + sum;
        }
        if (hx & BO_PKT_LAST) {
          double v0[32], b0[32];
          for (int l = 0; l < 32; ++l) bo_lane_target_load<MODE>(C, (unsigned)pkt[2 * l + 1] & 0x7FFFu, &v0[l], &b0[l]);
          const int g_round = G > 0 ? G : 1 << ((hx >> 4) & 7u);  // G = 0: lanes per target chosen per level, in the header
          for (int off = g_round >> 1; off > 0; off >>= 1) {
            double t[32];
            for (int l = 0; l < 32; ++l) t[l] = acc[w][l] + acc[w][l ^ off];
            for (int l = 0; l < 32; ++l) acc[w][l] = t[l];
          }
          for (int l = 0; l < 32; ++l) {
            const unsigned y = (unsigned)pkt[2 * l + 1];
            const int tgt = y & 0x7FFFu;
            if (tgt != 0x7FFF) bo_lane_finish<MODE>(C, tgt, (y & 0x8000u) != 0, acc[w][l], v0[l], b0[l]);
          }
        }
        pc[w] += 1;
        if (hx & BO_PKT_LEVEL_END) break;
      }
      more = more || pc[w] < n;
    }
    // A factorisation with a bad pivot is thrown away (the caller regularises and starts over), so the program is
    // abandoned at the end of the level in which the first one turns up: 44 % (C5) / 22 % (C4) of the way on average
    // (BO_FAC_ABANDON: only for the one-CTA-per-SM problems; on C3, 16 CTAs per SM and 3 % failing factorisations, the vote
    // at every level costs more than it saves: 893 k -> 849 k inst/s)
#ifdef BO_FAC_ABANDON
    if (MODE == 0 && C.ibuf[0] < BO_NK) break;
#endif
  }
#else
  const int warp = BO_TID >> 5, lane = BO_TID & 31;
  if (warp < W) {
    // every warp stream ends with BO_LP_PAD_PACKETS (>= BO_LP_DEPTH) padding packets, so the prefetch needs no bounds test:
    // one running pointer, loads at immediate offsets
    const int2* nxt = reinterpret_cast<const int2*>(C.tab + h[2]) + 32LL * h[4 + 2 * warp] + lane;
    const int n = h[5 + 2 * warp] / (PK + 1);  // packets of this warp
    bo_int2 buf[BO_LP_DEPTH][PK + 1];
    BO_UNROLL
    for (int d = 0; d < BO_LP_DEPTH; ++d) {
      BO_UNROLL
      for (int k = 0; k <= PK; ++k) {
        const int2 v = __ldg(nxt + 32 * k);
        buf[d][k] = bo_int2{v.x, v.y};
      }
      nxt += 32 * (PK + 1);
    }
    double acc = 0.0;
    bool bad_seen = false, abandon = false;
    for (int base = 0; base < n && !abandon; base += BO_LP_DEPTH) {
      BO_UNROLL
      for (int d = 0; d < BO_LP_DEPTH; ++d) {
        const int i = base + d;
        if (i >= n || abandon) break;
        bo_int2 cur[PK + 1];
        BO_UNROLL
        for (int k = 0; k <= PK; ++k) cur[k] = buf[d][k];
        BO_UNROLL
        for (int k = 0; k <= PK; ++k) {
          const int2 v = __ldg(nxt + 32 * k);
          buf[d][k] = bo_int2{v.x, v.y};
        }
        nxt += 32 * (PK + 1);
        const unsigned hx = (unsigned)cur[0].x, hy = (unsigned)cur[0].y;
        const int tgt = hy & 0x7FFFu;
        if (!(hx & BO_PKT_EMPTY)) {
          const double sum = bo_pkt_sum<MODE, PK>(C, cur + 1);
          acc = (hx & BO_PKT_FIRST) ? sum : acc + sum;
        }
        if (hx & BO_PKT_LAST) {
          double v0, b0;
          bo_lane_target_load<MODE>(C, tgt, &v0, &b0);
          if (G > 0) {
            BO_UNROLL
            for (int off = G >> 1; off > 0; off >>= 1) acc += bo_shfl_xor(acc, off);
          } else {  // lanes per target chosen per level (log2 in the header, warp-uniform)
            const int g_round = 1 << ((hx >> 4) & 7u);
            BO_UNROLL
            for (int off = 16; off > 0; off >>= 1)
              if (off < g_round) acc += bo_shfl_xor(acc, off);
          }
          if (tgt != 0x7FFF) bad_seen |= bo_lane_finish<MODE>(C, tgt, (hy & 0x8000u) != 0, acc, v0, b0);
        }
        if (hx & BO_PKT_LEVEL_END) {
          // factor: a bad pivot anywhere in this level ends the program for every warp (see the host branch above)
#ifdef BO_FAC_ABANDON
          if (MODE == 0) abandon = bo_bar_warps_or(W, bad_seen);
          else bo_bar_warps(W);
#else
          bo_bar_warps(W);
#endif
        }
      }
    }
  }
#endif
}

// In-place factorisation.  On exit vals[j] = 1/D(j), vals[n+e] = C(i,j) = L(i,j) D(j).
// Returns 0 ok, 1 = first bad pivot is in the x block (non-positive), 2 = in the y block (non-negative); "first" = the
// smallest column among the bad pivots of the levels done when the program was abandoned (it stops at the end of the level
// in which a bad pivot turns up: the factor is thrown away anyway).
BO_NOINLINE int bo_cta_factor(const bo_cta& C) {
  BO_PROF_BEGIN();
  if (BO_TID == 0) C.ibuf[0] = BO_NK;  // first bad pivot column (elimination order); BO_NK = none
  bo_sync();
  bo_lane_program<0, BO_FAC_G, BO_FAC_PK>(C, CT_PROG_FAC);
  bo_sync();
  const int badcol = C.ibuf[0];
  bo_sync();
  BO_PROF_END(3);
  BO_PROF_COUNT(5);
  if (badcol >= BO_NK) return 0;
  return (C.tab + C.tab[CT_SIGN])[badcol] > 0 ? 1 : 2;
}

// The factor program carries a right-hand side along (bo_coop.cpp): what is in bp when it starts comes out as L^-1 bp.
// bo_cta_load_rhs puts b there (before bo_cta_factor); after a successful factorisation bo_cta_ldl_solve_tail finishes
// K^-1 b with the scaling by 1/D and the backward program only.
BO_NOINLINE void bo_cta_load_rhs(const bo_cta& C, const double* b) {
  const int32_t* perm = C.tab + C.tab[CT_PERM];
  double* bp = BO_BP_P(C);
  BO_PAR(j, BO_NK) bp[j] = b[perm[j]];
  bo_sync();
}
BO_NOINLINE void bo_cta_ldl_solve_tail(const bo_cta& C, double* b) {
  BO_PROF_BEGIN();
  const int32_t* perm = C.tab + C.tab[CT_PERM];
  double* bp = BO_BP_P(C);
  BO_PAR(j, BO_NK) bp[j] *= BO_VALS_AT(C, j);
  bo_sync();
  bo_lane_program<2, BO_BWD_G, BO_BWD_PK>(C, CT_PROG_BWD);
  bo_sync();
  BO_PAR(j, BO_NK) b[perm[j]] = bp[j];
  bo_sync();
  BO_PROF_END(4);
  BO_PROF_COUNT(6);
}

// Solve K b = b (b indexed by original row: x then y) with the factorisation in vals.  The substitutions run
// on warp 0 (a level holds a couple of rows); ends with a CTA barrier.
BO_NOINLINE void bo_cta_ldl_solve(const bo_cta& C, double* b) {
  BO_PROF_BEGIN();
  const int32_t* perm = C.tab + C.tab[CT_PERM];
  double* bp = BO_BP_P(C);
  BO_PAR(j, BO_NK) bp[j] = b[perm[j]];
  bo_sync();
  bo_lane_program<1, BO_FWD_G, BO_FWD_PK>(C, CT_PROG_FWD);
  // the two programs may run on different numbers of warps: without this barrier a warp that has no part in the forward
  // program would start the backward one early (and meet named barrier 1 with another thread count: an illegal instruction)
  bo_sync();
  bo_lane_program<2, BO_BWD_G, BO_BWD_PK>(C, CT_PROG_BWD);
  bo_sync();
  BO_PAR(j, BO_NK) b[perm[j]] = bp[j];
  bo_sync();
  BO_PROF_END(4);
  BO_PROF_COUNT(6);
}

// ---------------------------------------------------------------------------------------------------
// the iteration (transcription of bo_ipm_reg.cuh's trip functions with CTA-parallel vector work)
// ---------------------------------------------------------------------------------------------------
BO_DEVICE void bo_cta_measures(const bo_cta& C, double f, const double* cE, const double* cI, const double* s, double mu,
                               double* phi, double* theta) {
  double v[2] = {0.0, 0.0};
  BO_PAR(j, BO_ME) v[0] += fabs(cE[j]);
  BO_PAR(i, BO_MI) {
    v[0] += fabs(cI[i] - s[i]);
    v[1] += log(s[i]);
  }
  const int op[2] = {BO_RED_SUM, BO_RED_SUM};
  bo_reduce<2>(v, op, BO_RED_P(C));
  *theta = v[0];
  *phi = f - mu * v[1];
}

BO_DEVICE void bo_cta_init(bo_cta_state& S, const bo_cta& C, const bo_solver_params& prm) {
  double* W = C.W;
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
  S.n_singular = 0;
  S.jac_degenerate = false;
  S.phase = BO_PH_EVAL;
  S.rhs_ready = false;
  S.trips = 0;
  S.resto = 0;
  S.n_resto = 0;
  bo_cta_pre(C);
  S.f = bo_cta_eval_fc(C, W + BO_OFF_X, W + BO_OFF_CE, W + BO_OFF_CI);
  BO_PAR(i, BO_MI) {
    const double ci = W[BO_OFF_CI + i];
    const double s = fmax(ci, 1e-2 * fmax(1.0, fabs(ci)));
    W[BO_OFF_S + i] = s;
    W[BO_OFF_Z + i] = S.mu / s;
  }
  BO_PAR(j, BO_ME) W[BO_OFF_Y + j] = 0.0;
  bo_sync();
}

// Restoration step data at x (bo_ipm_reg.cuh: bo_resto_prepare): weights / residuals of the violated inequality rows,
// no Hessian, no barrier, no multipliers.  Returns theta_r; ends with a CTA barrier.
BO_NOINLINE double bo_cta_resto_prepare(const bo_cta& C) {
  double* W = C.W;
  double v[1] = {0.0};
  BO_PAR(j, BO_ME) {
    const double c = W[BO_OFF_CE + j];
    W[BO_OFF_RE + j] = c;
    v[0] += c * c;
  }
  BO_PAR(i, BO_MI) {
    const double r = fmin(W[BO_OFF_CI + i], 0.0);
    W[BO_OFF_SIG + i] = r < 0.0 ? 1.0 : 0.0;
    W[BO_OFF_RI + i] = r;
    W[BO_OFF_Z + i] = 0.0;
    W[BO_OFF_S + i] = BO_INF;  // mu / s = 0, no fraction-to-the-boundary limit
    v[0] += r * r;
  }
  BO_PAR(c, BO_NX) W[BO_OFF_RD + c] = 0.0;
  BO_PAR(i, BO_NNZ_H) W[BO_OFF_H + i] = 0.0;
  const int op[1] = {BO_RED_SUM};
  bo_reduce<1>(v, op, BO_RED_P(C));
  bo_sync();
  return sqrt(v[0]);
}

// Step for the residuals (rE, rI) with the current factorisation: fills sol (dx, -dy), dx, ds; returns the
// fraction-to-the-boundary primal step length.
// ... its right-hand side into sol (needs S.rho, the residuals rE / rI, sigma, rd); ends with a CTA barrier
BO_NOINLINE void bo_cta_step_rhs(bo_cta_state& S, const bo_cta& C) {
  double* W = C.W;
  BO_PAR(i, BO_MI) W[BO_OFF_TV + i] = -(W[BO_OFF_Z + i] - S.mu / W[BO_OFF_S + i] + W[BO_OFF_SIG + i] * W[BO_OFF_RI + i]);
  BO_PAR(j, BO_ME) {
    W[BO_OFF_SOL + BO_NX + j] = -W[BO_OFF_RE + j];
    W[BO_OFF_T2 + j] = -S.rho * W[BO_OFF_RE + j];
  }
  bo_sync();
  BO_PAR(c, BO_NX) W[BO_OFF_SOL + c] = -W[BO_OFF_RD + c] + BO_JI_T(c, W + BO_OFF_TV) + BO_JE_T(c, W + BO_OFF_T2);
  bo_sync();
}
// `with_factor`: the right-hand side went through the factor program (bo_cta_load_rhs before bo_cta_factor): only the tail
// of the solve is left
BO_NOINLINE double bo_cta_step(bo_cta_state& S, const bo_cta& C, bool with_factor = false) {
  double* W = C.W;
  if (with_factor) {
    bo_cta_ldl_solve_tail(C, W + BO_OFF_SOL);
  } else {
    bo_cta_step_rhs(S, C);
    bo_cta_ldl_solve(C, W + BO_OFF_SOL);
  }
  const double undo = 1.0 / (1.0 - S.rho * S.dc);
  BO_PAR(j, BO_ME) W[BO_OFF_SOL + BO_NX + j] *= undo;
  BO_PAR(i, BO_NX) W[BO_OFF_DX + i] = W[BO_OFF_SOL + i];
  bo_sync();
  double v[1] = {1.0};
  BO_PAR(i, BO_MI) {
    const double ds = BO_JI_ROW(i, W + BO_OFF_DX) + W[BO_OFF_RI + i];
    W[BO_OFF_DS + i] = ds;
    if (ds < 0.0) v[0] = fmin(v[0], -S.tau * W[BO_OFF_S + i] / ds);
  }
  const int op[1] = {BO_RED_MIN};
  bo_reduce<1>(v, op, BO_RED_P(C));
  bo_sync();
  return v[0];
}

BO_DEVICE int bo_cta_trip_eval(bo_cta_state& S, const bo_cta& C, const bo_solver_params& prm) {
  const double kappa_eps = 10.0, kappa_mu = 0.2, tau_min = 0.99, s_max = 100.0;
  const double mu_min = prm.tol * 0.1;
  double* W = C.W;
  // trip budget: an instance over budget right after an accepted step is evaluated once more, so that f / err0 /
  // multipliers reported belong to the x returned (see bo_ipm_reg.cuh)
  const bool over = ++S.trips > prm.max_trips;
  if (over && S.phase != BO_PH_EVAL) return BO_ST_MAX_ITER;
  if (S.phase != BO_PH_EVAL) return -1;
  S.f = bo_cta_eval_kkt(C);
  if (S.resto > 0) {
    // restoration: next Levenberg-Marquardt step on the infeasibility from the fresh Jacobians
    if (!bo_isfinite(S.f)) return BO_ST_NUMERICAL;
    if (over || S.it >= prm.max_iter) return BO_ST_MAX_ITER;
    if (S.resto > BO_RESTO_MAX_IT) return BO_ST_LINE_SEARCH;
    S.thr = bo_cta_resto_prepare(C);
    S.dw = BO_RESTO_ZETA;
    S.dc = BO_RESTO_DC;
    S.first_singular = false;
    S.attempt = 0;
    S.heavy = 0;
    S.ls_mode = false;
    S.phase = BO_PH_FACTOR;
    S.rhs_ready = false;
    return -1;
  }
  if (S.recalc_y && BO_ME > 0 && !over) {
    // least-squares multiplier estimate after a regularised step (see bo_ipm_reg.cuh)
    S.recalc_y = false;
    S.ls_mode = true;
    BO_PAR(i, BO_NNZ_H) W[BO_OFF_H + i] = 0.0;
    BO_PAR(i, BO_MI) W[BO_OFF_SIG + i] = 0.0;
    bo_sync();
    S.dw = 1.0;
    S.dc = 1e-10;
    S.phase = BO_PH_FACTOR;
    S.rhs_ready = false;
    return -1;
  }
  // residuals and the scaled optimality error
  double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};  // e_dual, e_prim, e_comp0 (max) ; sum |y|, sum |z|
  BO_PAR(c, BO_NX) {
    const double rd = W[BO_OFF_G + c] - BO_JE_T(c, W + BO_OFF_Y) - BO_JI_T(c, W + BO_OFF_Z);
    W[BO_OFF_RD + c] = rd;
    v[0] = fmax(v[0], fabs(rd));
  }
  BO_PAR(j, BO_ME) {
    v[1] = fmax(v[1], fabs(W[BO_OFF_CE + j]));
    v[3] += fabs(W[BO_OFF_Y + j]);
  }
  BO_PAR(i, BO_MI) {
    v[1] = fmax(v[1], fabs(W[BO_OFF_CI + i] - W[BO_OFF_S + i]));
    v[2] = fmax(v[2], W[BO_OFF_S + i] * W[BO_OFF_Z + i]);
    v[4] += fabs(W[BO_OFF_Z + i]);
  }
  {
    const int op[5] = {BO_RED_MAX, BO_RED_MAX, BO_RED_MAX, BO_RED_SUM, BO_RED_SUM};
    bo_reduce<5>(v, op, BO_RED_P(C));
  }
  const double e_dual = v[0], e_prim = v[1], e_comp0 = v[2], sum_z = v[4], sum_mult = v[3] + v[4];
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
  // barrier parameter update (monotone Fiacco-McCormick); resets the filter
  if (BO_MI > 0) {
    for (int rep = 0; rep < 8; ++rep) {
      double w1[1] = {0.0};
      BO_PAR(i, BO_MI) w1[0] = fmax(w1[0], fabs(W[BO_OFF_S + i] * W[BO_OFF_Z + i] - S.mu));
      const int op[1] = {BO_RED_MAX};
      bo_reduce<1>(w1, op, BO_RED_P(C));
      const double err_mu = fmax(fmax(e_dual / s_d, e_prim), w1[0] / s_c);
      if (err_mu <= kappa_eps * S.mu && S.mu > mu_min) {
        S.mu = fmax(mu_min, fmin(kappa_mu * S.mu, S.mu * sqrt(S.mu)));
        S.nf = 0;
      } else {
        break;
      }
    }
  }
  S.tau = fmax(tau_min, 1.0 - S.mu);
  BO_PAR(i, BO_MI) W[BO_OFF_SIG + i] = W[BO_OFF_Z + i] / W[BO_OFF_S + i];
  bo_cta_measures(C, S.f, W + BO_OFF_CE, W + BO_OFF_CI, W + BO_OFF_S, S.mu, &S.phi0, &S.theta0);
  if (S.it == 0) {
    S.theta_max = 1e4 * fmax(1.0, S.theta0);
    S.theta_min = 1e-4 * fmax(1.0, S.theta0);
  }
  S.dw = 0.0;
  // IPOPT's degeneracy heuristic (PDPerturbationHandler): once the constraint Jacobian has been found rank deficient
  // in three consecutive iterations, the constraint block is perturbed from the first attempt on
  S.dc = S.jac_degenerate ? BO_DC_SCALE * sqrt(sqrt(S.mu)) : 0.0;
  S.first_singular = false;
  S.attempt = 0;
  S.heavy = 0;
  S.ls_mode = false;
  S.phase = BO_PH_FACTOR;
  S.rhs_ready = false;
  bo_sync();
  return -1;
}

BO_DEVICE int bo_cta_trip_factor(bo_cta_state& S, const bo_cta& C, const bo_solver_params& prm) {
  double* W = C.W;
  if (S.phase != BO_PH_FACTOR) return -1;
  const double rho = S.ls_mode ? 0.0 : (S.resto > 0 ? 1.0 : BO_STATIC_RHO);
  S.rho = rho;
  // the right-hand side of the solve that follows goes through the factor program (forward substitution for free); it does
  // not depend on the regularisation, so inertia retries reuse it
  if (!S.rhs_ready) {
    if (S.ls_mode) {
      BO_PAR(c, BO_NX) W[BO_OFF_SOL + c] = W[BO_OFF_G + c] - BO_JI_T(c, W + BO_OFF_Z);
      BO_PAR(j, BO_ME) W[BO_OFF_SOL + BO_NX + j] = 0.0;
      bo_sync();
    } else {
      if (S.resto == 0) {
        BO_PAR(j, BO_ME) W[BO_OFF_RE + j] = W[BO_OFF_CE + j];
        BO_PAR(i, BO_MI) W[BO_OFF_RI + i] = W[BO_OFF_CI + i] - W[BO_OFF_S + i];
        bo_sync();
      }
      bo_cta_step_rhs(S, C);
    }
    S.rhs_ready = true;
  }
  bo_cta_assemble(C, rho, S.dw, S.dc / (1.0 - rho * S.dc));
  bo_cta_load_rhs(C, W + BO_OFF_SOL);
  const int bad = bo_cta_factor(C);
  const int inertia = bad == 0 ? 0 : (bad == 1 ? 1 : -1);
  if (S.ls_mode) {
    if (inertia == 0) {
      S.rhs_ready = false;
      bo_cta_ldl_solve_tail(C, W + BO_OFF_SOL);
      double v[1] = {0.0};
      BO_PAR(j, BO_ME) v[0] = fmax(v[0], bo_isfinite(W[BO_OFF_SOL + BO_NX + j]) ? 0.0 : 1.0);
      const int op[1] = {BO_RED_MAX};
      bo_reduce<1>(v, op, BO_RED_P(C));
      if (v[0] == 0.0) {
        BO_PAR(j, BO_ME) W[BO_OFF_Y + j] = W[BO_OFF_SOL + BO_NX + j];
        // One step of iterative refinement towards the UNregularised least-squares multipliers: the factorised
        // system carries -dc on the constraint block, which scales every component of y by s^2 / (s^2 + dc)
        // (s: singular value of JE) -- a relative error of dc / s^2 ~ 1e-8 that would otherwise sit in the dual
        // infeasibility as a floor just above tol (C4 stalled at 2.7e-8 and ended "acceptable").  Only near
        // convergence: far from it the extra accuracy buys nothing and costs a solve per iteration.
        if (S.err0 < BO_REFINE_BELOW) {
        BO_PAR(c, BO_NX) W[BO_OFF_DX0 + c] = W[BO_OFF_SOL + c];
        bo_sync();
        BO_PAR(c, BO_NX)
          W[BO_OFF_SOL + c] = (W[BO_OFF_G + c] - BO_JI_T(c, W + BO_OFF_Z)) - (S.dw * W[BO_OFF_DX0 + c] + BO_JE_T(c, W + BO_OFF_Y));
        BO_PAR(j, BO_ME)
          W[BO_OFF_SOL + BO_NX + j] = -bo_gather(C.tab + C.tab[CT_JE_RPTR], C.tab + C.tab[CT_JE_RENT], j, W + BO_OFF_JE, W + BO_OFF_DX0);
        bo_sync();
        bo_cta_ldl_solve(C, W + BO_OFF_SOL);
        BO_PAR(j, BO_ME) {
          const double d = W[BO_OFF_SOL + BO_NX + j];
          if (bo_isfinite(d)) W[BO_OFF_Y + j] += d;
        }
        }
      }
      bo_sync();
    }
    S.ls_mode = false;
    S.phase = BO_PH_EVAL;
    return -1;
  }
#ifdef BO_HOST_TRACE
  if (inertia != 0) printf("     inertia %d at dw %.3e dc %.3e\n", inertia, S.dw, S.dc);
#endif
  if (inertia != 0) {
    if (inertia < 0 && BO_ME > 0 && S.dc == 0.0) {
      S.dc = BO_DC_SCALE * sqrt(sqrt(S.mu));
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
  S.rhs_ready = false;  // the solve below overwrites sol
  const double a_p = bo_cta_step(S, C, true);
  BO_PAR(j, BO_ME) W[BO_OFF_YST + j] = -W[BO_OFF_SOL + BO_NX + j];
  double v[3] = {0.0, 0.0, 0.0};  // g'dx, sum ds/s, max |dx|
  BO_PAR(i, BO_NX) {
    v[0] += W[BO_OFF_G + i] * W[BO_OFF_DX + i];
    v[2] = fmax(v[2], fabs(W[BO_OFF_DX + i]));
  }
  BO_PAR(i, BO_MI) v[1] += W[BO_OFF_DS + i] / W[BO_OFF_S + i];
  {
    const int op[3] = {BO_RED_SUM, BO_RED_SUM, BO_RED_MAX};
    bo_reduce<3>(v, op, BO_RED_P(C));
  }
  S.dphi = v[0] - S.mu * v[1];
  const double dxn = v[2];
  S.a = a_p;
  if (prm.max_step > 0.0 && S.a * dxn > prm.max_step) S.a = prm.max_step / dxn;
  S.a_trial = S.a;
  S.ls = 0;
  S.soc = 0;
  S.phase = BO_PH_TRIAL;
  bo_sync();
  return -1;
}

BO_DEVICE int bo_cta_trip_trial(bo_cta_state& S, const bo_cta& C, const bo_solver_params& prm) {
  const double kappa_sigma = 1e10, gamma_theta = 1e-5, gamma_phi = 1e-5, eta_phi = 1e-8;
  const double s_phi = 2.3, s_theta = 1.1, kappa_soc = 0.99;
  double* W = C.W;
  if (S.phase != BO_PH_TRIAL) return -1;
  BO_PAR(i, BO_NX) W[BO_OFF_XT + i] = W[BO_OFF_X + i] + S.a_trial * W[BO_OFF_DX + i];
  BO_PAR(i, BO_MI) W[BO_OFF_ST + i] = W[BO_OFF_S + i] + S.a_trial * W[BO_OFF_DS + i];
  bo_sync();
  const double ft = bo_cta_eval_fc(C, W + BO_OFF_XT, W + BO_OFF_CET, W + BO_OFF_CIT);
  if (S.resto > 0) {
    // restoration trial point: Armijo on theta_r
    double v[1] = {0.0};
    BO_PAR(j, BO_ME) v[0] += W[BO_OFF_CET + j] * W[BO_OFF_CET + j];
    BO_PAR(i, BO_MI) {
      const double r = fmin(W[BO_OFF_CIT + i], 0.0);
      v[0] += r * r;
    }
    const int op[1] = {BO_RED_SUM};
    bo_reduce<1>(v, op, BO_RED_P(C));
    const double thr_t = sqrt(v[0]);
#ifdef BO_HOST_TRACE
    printf("     resto %d ls %d a %.3e theta_r %.3e -> %.3e (entry %.3e)\n", S.resto, S.ls, S.a_trial, S.thr, thr_t, S.thr0);
#endif
    if (bo_isfinite(thr_t) && thr_t <= (1.0 - 1e-4 * S.a_trial) * S.thr) {
      BO_PAR(i, BO_NX) W[BO_OFF_X + i] = W[BO_OFF_XT + i];
      S.it += 1;
      if (thr_t <= fmax(BO_RESTO_KAPPA * S.thr0, 1e-10)) {
        // feasible enough: back to the regular iteration from here, multipliers and filter start afresh
        S.f = ft;
        BO_PAR(i, BO_MI) {
          const double ci = W[BO_OFF_CIT + i];
          const double s = fmax(ci, 1e-2 * fmax(1.0, fabs(ci)));
          W[BO_OFF_S + i] = s;
          W[BO_OFF_Z + i] = S.mu / s;
        }
        BO_PAR(j, BO_ME) W[BO_OFF_Y + j] = 0.0;
        S.resto = 0;
        S.nf = 0;
        S.dw_last = 0.0;
        S.recalc_y = BO_ME > 0;  // least-squares equality multipliers at the new point
      } else {
        S.resto += 1;
      }
      S.phase = BO_PH_EVAL;
      bo_sync();
      return -1;
    }
    if (S.ls >= 2 && S.attempt < 4) {  // Levenberg-Marquardt: more damping, new direction
      S.attempt += 1;
      S.dw *= 100.0;
      S.phase = BO_PH_FACTOR;
      S.rhs_ready = false;
      return -1;
    }
    S.a *= 0.5;
    S.a_trial = S.a;
    S.ls += 1;
    if (S.ls >= 24 || S.a < 1e-10) return BO_ST_LINE_SEARCH;  // stationary point of the infeasibility
    return -1;
  }
  double phit, thetat;
  bo_cta_measures(C, ft, W + BO_OFF_CET, W + BO_OFF_CIT, W + BO_OFF_ST, S.mu, &phit, &thetat);
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
    if (!(ftype && armijo)) {
      int slot = S.nf;
      if (S.nf < BO_NFILTER) {
        ++S.nf;
      } else {
        slot = 0;
        for (int j = 1; j < BO_NFILTER; ++j)
          if (S.fth[j] > S.fth[slot]) slot = j;
      }
      S.fth[slot] = (1.0 - gamma_theta) * S.theta0;
      S.fph[slot] = S.phi0 - gamma_phi * S.theta0;
    }
    // dual step with its own fraction-to-the-boundary rule (dz kept in TV)
    double v[1] = {1.0};
    BO_PAR(i, BO_MI) {
      const double dz = -W[BO_OFF_Z + i] + S.mu / W[BO_OFF_S + i] - W[BO_OFF_SIG + i] * W[BO_OFF_DS + i];
      W[BO_OFF_TV + i] = dz;
      if (dz < 0.0) v[0] = fmin(v[0], -S.tau * W[BO_OFF_Z + i] / dz);
    }
    const int op[1] = {BO_RED_MIN};
    bo_reduce<1>(v, op, BO_RED_P(C));
    const double a_d = v[0];
    BO_PAR(i, BO_NX) W[BO_OFF_X + i] = W[BO_OFF_XT + i];
    BO_PAR(i, BO_MI) {
      const double s = fmax(W[BO_OFF_ST + i], W[BO_OFF_CIT + i]);
      W[BO_OFF_S + i] = s;
      double z = W[BO_OFF_Z + i] + a_d * W[BO_OFF_TV + i];
      z = fmax(fmin(z, kappa_sigma * S.mu / s), S.mu / (kappa_sigma * s));
      W[BO_OFF_Z + i] = z;
    }
    BO_PAR(j, BO_ME) W[BO_OFF_Y + j] += S.a * W[BO_OFF_YST + j];
    // Least-squares multiplier re-estimate after a regularised step (see bo_ipm_reg.cuh).  With BO_RECALC_SKIP_DEGENERATE
    // (bo_coop.cpp defines it for KKT systems of 600 rows and more) not when dc > 0 only because the degeneracy heuristic has
    // switched it on for good (jac_degenerate: a structurally rank-deficient Jacobian, e.g. the four quaternion equalities
    // per knot of figure_eight_plan.py, rank 3): there the re-estimate would run on EVERY iteration -- one more
    // factorisation and one or two more substitutions each -- where IPOPT (recalc_y = no) does none.  Measured: C4 (1250
    // rows; 32 instances, host build) 109 -> 77 factorisations and 144 -> 106 substitutions per instance for 8 % more
    // iterations, same minimisers, 1037 -> 1685 inst/s on B200; the joint-space planner (434 rows) 5.4 -> 7.4 iterations
    // and 51.0 k -> 48.6 k inst/s, i.e. below that size the re-estimate pays for itself and stays (when y enters the Hessian).
    // BO_RECALC_MODE (experiments): 0 never, 1 only after dw > 0, 2 also on every dc > 0, 3 skip under jac_degenerate.
#if defined(BO_RECALC_MODE) && BO_RECALC_MODE == 0
    S.recalc_y = false;
#elif defined(BO_RECALC_MODE) && BO_RECALC_MODE == 1
    S.recalc_y = S.dw > 0.0;
#elif defined(BO_RECALC_DC_ONLY)  /* y does not enter the Hessian (linear equalities): only the rank-deficient case needs it */
    // ... and there the estimate buys little once dc is on for good, whatever the size (C3 from the zero seed: 72 ms per
    // 16384 without it, 79 ms with it)
#if !(defined(BO_RECALC_MODE) && BO_RECALC_MODE == 2)
    S.recalc_y = S.dc > 0.0 && !S.jac_degenerate;
#else
    S.recalc_y = S.dc > 0.0;
#endif
#else
#if (defined(BO_RECALC_SKIP_DEGENERATE) && !(defined(BO_RECALC_MODE) && BO_RECALC_MODE == 2)) || (defined(BO_RECALC_MODE) && BO_RECALC_MODE == 3)
    S.recalc_y = S.dw > 0.0 || (S.dc > 0.0 && !S.jac_degenerate);
#else
    S.recalc_y = S.dw > 0.0 || S.dc > 0.0;
#endif
#endif
    S.it += 1;
    S.phase = BO_PH_EVAL;
    bo_sync();
    return -1;
  }
  bool try_soc = false;
  if (S.soc == 0) {
    if (S.ls == 0 && finite && thetat >= S.theta0 && (BO_ME + BO_MI) > 0) {
      BO_PAR(i, BO_NX) W[BO_OFF_DX0 + i] = W[BO_OFF_DX + i];
      BO_PAR(i, BO_MI) W[BO_OFF_DS0 + i] = W[BO_OFF_DS + i];
      BO_PAR(j, BO_ME) W[BO_OFF_RE + j] = S.a * W[BO_OFF_CE + j] + W[BO_OFF_CET + j];
      BO_PAR(i, BO_MI) W[BO_OFF_RI + i] = S.a * (W[BO_OFF_CI + i] - W[BO_OFF_S + i]) + (W[BO_OFF_CIT + i] - W[BO_OFF_ST + i]);
      S.th_soc = thetat;
      try_soc = true;
    }
  } else if (S.soc < 4 && finite && thetat <= kappa_soc * S.th_soc) {
    BO_PAR(j, BO_ME) W[BO_OFF_RE + j] = S.a_trial * W[BO_OFF_RE + j] + W[BO_OFF_CET + j];
    BO_PAR(i, BO_MI) W[BO_OFF_RI + i] = S.a_trial * W[BO_OFF_RI + i] + (W[BO_OFF_CIT + i] - W[BO_OFF_ST + i]);
    S.th_soc = thetat;
    try_soc = true;
  }
  if (try_soc) {
    bo_sync();
    S.a_trial = bo_cta_step(S, C);
    S.soc += 1;
    return -1;
  }
  if (S.soc > 0) {
    BO_PAR(i, BO_NX) W[BO_OFF_DX + i] = W[BO_OFF_DX0 + i];
    BO_PAR(i, BO_MI) W[BO_OFF_DS + i] = W[BO_OFF_DS0 + i];
    bo_sync();
    S.soc = 0;
  }
  S.a *= 0.5;
  S.a_trial = S.a;
  S.ls += 1;
  if ((S.ls >= BO_LS_MAX || S.a < BO_ALPHA_MIN) && S.theta0 > 1e-7 * fmax(1.0, S.theta_min * 1e4) && S.n_resto < BO_RESTO_MAX_PHASES) {
    // no acceptable step at an infeasible point: feasibility restoration (the evaluation at x is still valid)
    S.n_resto += 1;
    S.resto = 1;
    S.thr0 = S.thr = bo_cta_resto_prepare(C);
    S.dw = BO_RESTO_ZETA;
    S.dc = BO_RESTO_DC;
    S.attempt = 0;
    S.heavy = 0;
    S.phase = BO_PH_FACTOR;
    S.rhs_ready = false;
    return -1;
  }
  if (S.ls >= BO_LS_MAX || S.a < BO_ALPHA_MIN) {
    if (++S.heavy >= BO_HEAVY_MAX) return S.err0 <= prm.acceptable_tol ? BO_ST_ACCEPTABLE : BO_ST_LINE_SEARCH;  // IPOPT: a failed step at an acceptable point ends "solved to acceptable level"
    S.dw = fmax(S.dw * 100.0, 1.0);
    S.phase = BO_PH_FACTOR;
    S.rhs_ready = false;
  }
  return -1;
}

// One instance, start to finish.  W[P], W[X] hold the parameters and the seed on entry.
BO_DEVICE int bo_cta_solve(bo_cta_state& S, const bo_cta& C, const bo_solver_params& prm) {
  bo_cta_init(S, C, prm);
  int status;
  do {
    status = bo_cta_trip_eval(S, C, prm);
    if (status < 0) status = bo_cta_trip_factor(S, C, prm);
    if (status < 0) status = bo_cta_trip_trial(S, C, prm);
  } while (status < 0);
  return status;
}

#ifndef BO_HOST_SIM
// Persistent CTAs, instances fetched from a global counter.  Same signature as the thread-per-instance
// kernel; prm.scratch = per-CTA vector workspace ([gridDim.x][BO_SCRATCH_DOUBLES]).
extern "C" __global__ void __launch_bounds__(BO_TPB, BO_MIN_CTAS)
bo_solve_kernel(long long B, const double* __restrict__ p_all, const double* __restrict__ x0_all,
                double* __restrict__ x_all, double* __restrict__ lam_all, double* __restrict__ f_all,
                int* __restrict__ status_all, int* __restrict__ iters_all, double* __restrict__ kkt_all,
                unsigned long long* __restrict__ work_counter, const bo_solver_params prm) {
  bo_cta C;
  C.red = bo_smem + BO_SM_RED;
  C.ibuf = reinterpret_cast<int*>(bo_smem + BO_SM_INTS);
  C.prof = nullptr;
  C.vals = bo_smem + BO_SM_VALS;
  C.bp = bo_smem + BO_SM_BP;
  C.wkkt = bo_smem + BO_SM_VALS;
  C.wfc = bo_smem + BO_SM_FREE;
#ifdef BO_PROFILE
  __shared__ long long bo_prof[8];
  C.prof = bo_prof;
#endif
#ifdef BO_W_IN_SMEM
  // small problems (C3: 27 KB of vectors per instance): the whole per-instance state stays on chip.  Round 1 kept it in a
  // per-CTA slice of global memory "resident in L2": 5.6 GB of DRAM traffic per launch for 32 MB of algorithmic I/O
  // (profiles/r01_coop_c3_details.txt) and an L2 round trip under every vector operation.
  C.W = bo_smem + BO_SMEM_DOUBLES;
#else
  C.W = prm.scratch + (long long)blockIdx.x * BO_SCRATCH_DOUBLES;
#endif
  C.tab = prm.ldl_tab;
  C.dtab = prm.dtab;
  bo_cta_state S;
  while (true) {
    if (threadIdx.x == 0) {
      const unsigned long long b = atomicAdd(work_counter, 1ULL);
      C.ibuf[1] = b < (unsigned long long)B ? (int)b : -1;
    }
    __syncthreads();
    const long long b = C.ibuf[1];
    __syncthreads();
    if (b < 0) break;
    BO_PAR(i, BO_NP) C.W[BO_OFF_P + i] = p_all[b * BO_NP + i];
    BO_PAR(i, BO_NX) C.W[BO_OFF_X + i] = x0_all ? x0_all[b * BO_NX + i] : 0.0;
    __syncthreads();
#ifdef BO_PROFILE
    if (threadIdx.x < 8) C.prof[threadIdx.x] = 0;
    const long long bo_t_start = clock64();
#endif
    const int status = bo_cta_solve(S, C, prm);
    __syncthreads();
    BO_PAR(i, BO_NX) x_all[b * BO_NX + i] = C.W[BO_OFF_X + i];
#ifdef BO_PROFILE
    // profiling build: the first 8 doubles of x are replaced by the cycle counters
    // (0 kkt tape, 1 f/c tape, 2 assembly, 3 factorisation, 4 solves, 5 #factorisations, 6 #solves, 7 whole instance)
    __syncthreads();
    if (threadIdx.x == 0) {
      C.prof[7] = clock64() - bo_t_start;
      for (int k = 0; k < 8 && k < BO_NX; ++k) x_all[b * BO_NX + k] = (double)C.prof[k];
    }
#endif
    if (lam_all) {
      BO_PAR(j, BO_ME) lam_all[b * (BO_ME + BO_MI) + j] = C.W[BO_OFF_Y + j];
      BO_PAR(i, BO_MI) lam_all[b * (BO_ME + BO_MI) + BO_ME + i] = C.W[BO_OFF_Z + i];
    }
    if (threadIdx.x == 0) {
      if (f_all) f_all[b] = S.f;
      if (status_all) status_all[b] = status;
      if (iters_all) iters_all[b] = S.it;
      if (kkt_all) kkt_all[b] = S.err0;
    }
    __syncthreads();
  }
}
#endif
