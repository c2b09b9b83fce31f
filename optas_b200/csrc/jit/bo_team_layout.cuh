// bo_team_layout.cuh -- shared-memory layout and tuning constants of the team tier (bo_ipm_team.cuh); included by the
// generated tape slices, which address the per-instance state directly.
#pragma once
#include "bo_common.cuh"

#define BO_NK (BO_NX + BO_ME)
#define BO_KSZ ((BO_NK * (BO_NK + 1)) / 2)
#define BO_KIDX(i, j) (((i) * ((i) + 1)) / 2 + (j)) /* packed lower triangle, i >= j */
#define BO_DIM(n) ((n) > 0 ? (n) : 1)

#ifndef BO_DC_SCALE
#define BO_DC_SCALE 1e-8
#endif
#ifndef BO_STATIC_RHO /* defined by the generated prelude: the same value is baked into the KX tape outputs */
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
#define BO_PH_DONE 4 /* finished in M1; results are written out at the end of the trip */
#define BO_PH_FRESH 5 /* parameters and seed of a new instance are staged; the team's quad has not picked it up yet */

// ---- shared-memory layout: offsets in "elements" (one element = BO_LS doubles, one per lane) ----
#define BO_OFF_P 0
#define BO_OFF_X (BO_OFF_P + BO_NP)
#define BO_OFF_XT (BO_OFF_X + BO_NX)      /* trial point x + a dx (becomes x when the step is accepted) */
#define BO_OFF_SH (BO_OFF_XT + BO_NX)     /* sin / cos of XT (2k, 2k + 1), shared by all slices */
#define BO_OFF_S (BO_OFF_SH + 2 * BO_NX)
#define BO_OFF_Z (BO_OFF_S + BO_MI)
#define BO_OFF_RS (BO_OFF_Z + BO_MI)      /* 1 / s */
#define BO_OFF_SIG (BO_OFF_RS + BO_MI)    /* z / s */
#define BO_OFF_Y (BO_OFF_SIG + BO_MI)
#define BO_OFF_SN (BO_OFF_Y + BO_ME)      /* slacks of the trial point after the slack reset, and their reciprocals */
#define BO_OFF_RSN (BO_OFF_SN + BO_MI)
#define BO_OFF_F0 (BO_OFF_RSN + BO_MI)    /* f at x */
#define BO_OFF_G (BO_OFF_F0 + 1)
#define BO_OFF_CE (BO_OFF_G + BO_NX)
#define BO_OFF_CI (BO_OFF_CE + BO_ME)
#define BO_OFF_JE (BO_OFF_CI + BO_MI)
#define BO_OFF_JI (BO_OFF_JE + BO_NNZ_JE)
#define BO_OFF_KX (BO_OFF_JI + BO_NNZ_JI)  /* packed lower triangle of H + JI' diag(sigma) JI + rho JE'JE */
#define BO_OFF_RD (BO_OFF_KX + (BO_NX * (BO_NX + 1)) / 2)
#define BO_OFF_LD (BO_OFF_RD + BO_NX)     /* packed factor: 1/D on the diagonal, unit-lower L below */
#define BO_OFF_SOL (BO_OFF_LD + BO_KSZ)   /* right-hand side / solution of the KKT solves */
#define BO_OFF_DX (BO_OFF_SOL + BO_NK)
#define BO_OFF_DS (BO_OFF_DX + BO_NX)
#define BO_OFF_YST (BO_OFF_DS + BO_MI)
#define BO_OFF_DX0 (BO_OFF_YST + BO_ME)
#define BO_OFF_DS0 (BO_OFF_DX0 + BO_NX)
#define BO_OFF_RE (BO_OFF_DS0 + BO_MI)
#define BO_OFF_RI (BO_OFF_RE + BO_ME)
#define BO_OFF_FT (BO_OFF_RI + BO_MI)     /* f at the trial point */
#define BO_OFF_CET (BO_OFF_FT + 1)
#define BO_OFF_CIT (BO_OFF_CET + BO_ME)
#define BO_OFF_PART (BO_OFF_CIT + BO_MI)  /* per role, from the f / c slices: sum log(s_trial), sum |c_trial|, sum log(s_new);
                                             from the KKT slices: max |c|, sum |c|, min s z, max s z, sum |z| */
#define BO_OFF_AT (BO_OFF_PART + 5 * BO_G) /* trial step length, published by the master */
#define BO_OFF_FTH (BO_OFF_AT + 1)        /* the filter */
#define BO_OFF_FPH (BO_OFF_FTH + BO_NFILTER)
#define BO_SM_ELEMS (BO_OFF_FPH + BO_NFILTER)

#ifdef BO_HOST_SIM
#define BO_LS 1
#else
#define BO_LS 32
#endif
#define SM(off, i) sm[((off) + (i)) * BO_LS]
#define SMP(off) (sm + (off) * BO_LS)

