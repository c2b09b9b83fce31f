// bo_common.cuh -- shared prelude of every JIT-compiled kernel of libb200optas (sm_100a, NVRTC).
//
// The same text also compiles as plain host C++ when BO_HOST_SIM is defined: that build is a
// TEST HARNESS ONLY (tools/hostsim), used on GPU-less CI to exercise the solver logic; it is
// never linked into, or reachable from, libb200optas.so.
#pragma once

#ifdef BO_HOST_SIM
#include <cmath>
#include <cstdint>
#include <limits>
#define BO_DEVICE static inline
#define BO_NOINLINE static __attribute__((noinline))
#define BO_RESTRICT __restrict__
#define BO_UNROLL
#define BO_NOUNROLL
#define BO_CONSTANT static const
#define BO_INF (std::numeric_limits<double>::infinity())
#define BO_NAN (std::numeric_limits<double>::quiet_NaN())
static inline void bo_sincos(double a, double* s, double* c) { *s = std::sin(a); *c = std::cos(a); }
static inline bool bo_isfinite(double v) { return std::isfinite(v); }
using std::fabs; using std::fmax; using std::fmin; using std::sqrt; using std::log; using std::pow;
using std::floor; using std::ceil; using std::exp; using std::atan2; using std::sin; using std::cos; using std::tan;
using std::asin; using std::acos; using std::atan; using std::sinh; using std::cosh; using std::tanh;
#else
#define BO_DEVICE __device__ __forceinline__
#define BO_NOINLINE __device__ __noinline__
#define BO_RESTRICT __restrict__
#define BO_UNROLL _Pragma("unroll")
#define BO_NOUNROLL _Pragma("unroll 1") /* loops whose body is a long libm expansion (div, log): keep one copy */
#define BO_CONSTANT __constant__
#define BO_INF (__longlong_as_double(0x7ff0000000000000LL))
#define BO_NAN (__longlong_as_double(0x7ff8000000000000LL))
typedef int int32_t;
typedef long long int64_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
__device__ __forceinline__ void bo_sincos(double a, double* s, double* c) { sincos(a, s, c); }
__device__ __forceinline__ bool bo_isfinite(double v) { return isfinite(v); }
#endif

struct alignas(16) bo_int4 { int x, y, z, w; };
struct alignas(8) bo_int2 { int x, y; };

BO_DEVICE double bo_sign(double a) { return (double)((a > 0.0) - (a < 0.0)); }
BO_DEVICE double bo_sq(double a) { return a * a; }

// Parameters of one bo_solve launch (passed by value as a kernel argument).
struct bo_solver_params {
  int32_t max_iter;
  double tol;
  double acceptable_tol;
  double mu_init;
  double max_step;  // <= 0: unlimited
  int32_t max_trips; // budget of solver trips per instance (bounds the tail of a batch)
  const int32_t* ldl_tab;  // int tables (bo_sparse.cpp): sparse LDL' plan; in large mode also structure + tapes
  const double* dtab;      // double table: constants of the interpreted tapes (large mode)
  double* scratch;         // large mode: factor values of all lanes, [element][lane]
  long long scratch_stride;  // = number of lanes of the launch
};

// per-instance status codes (mirror bo_instance_status in include/b200optas.h)
#define BO_ST_CONVERGED 0
#define BO_ST_ACCEPTABLE 1
#define BO_ST_MAX_ITER 2
#define BO_ST_LINE_SEARCH 3
#define BO_ST_NUMERICAL 4

// IPOPT's Jacobian-degeneracy heuristic (PDPerturbationHandler) counts an iteration as "singular" when the KKT matrix
// with dc = 0 was reported singular and dc > 0 cured it -- at whatever Hessian perturbation dw that happened, not only on
// the unperturbed first attempt (1: any attempt; 0: round-1 behaviour, first attempt only).  Three such iterations in a
// row switch dc on from the first attempt: in a non-convex region (dw > 0 every iteration) that saves the
// (dw, 0) factorisation of every later iteration without changing the system that is finally solved.
#ifndef BO_SINGULAR_ANY_ATTEMPT
#define BO_SINGULAR_ANY_ATTEMPT 1
#endif
