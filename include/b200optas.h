/* b200optas.h -- C ABI of libb200optas.so, the B200 (sm_100a) batched NLP/QP solver back-end
 * that sits behind the OpTaS solver interface.
 *
 * The reference (cmower/optas) has no C ABI: its solver plug-in boundary is the Python ABC
 * optas/solver.py:61-315 (`Solver`), whose concrete CasADiSolver crosses into native code at
 *     optas/solver.py:382   self._solver = cs.nlpsol/qpsol("solver", name, {x,p,f,g}, opts)   (create)
 *     optas/solver.py:395   self._solution = self._solver(x0=, p=, lbg=, ubg=)                (solve)
 *     optas/solver.py:396   self._stats = self._solver.stats()                                (status)
 * and whose model functions cross at
 *     optas/models.py:786-787   cs.Function(label,[q],[expr]).map(n)                          (batched FK eval)
 * Each entry point below replaces exactly one of those crossings; the comment on each says
 * which.  The binding a maintainer of the reference would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - every function returns BO_OK (0) or a negative bo_status; none throws, none calls back
 *     into the host language; bo_last_error() gives the message for the calling thread.
 *   - all numeric data is IEEE binary64, row-major with the INSTANCE index slowest:
 *     p[B][np], x[B][nx] ... (one OpTaS problem instance per row).
 *   - data pointers may be host or device memory (detected with cudaPointerGetAttributes);
 *     a call whose buffers are all device memory is asynchronous on `stream`; a call with any
 *     host buffer returns after the results are on the host.  PAGE-LOCKED host buffers are used in place by the
 *     kernel (zero copy: an instance reads its row when it starts and writes its results when it ends, so the PCIe
 *     traffic overlaps the other instances' iterations; B200OPTAS_ZERO_COPY=0 in the environment restores staged
 *     copies); pageable host buffers are staged through device buffers of the handle.
 *   - buffers are BORROWED for the duration of the call; nothing is retained.
 *   - a handle is bound to the CUDA device current at creation; handles are not thread-safe,
 *     distinct handles may be used from distinct threads / processes (one per GPU).
 *   - there is NO CPU fallback: without a usable sm_100-class device the solve/eval calls
 *     fail with BO_ERR_NO_DEVICE.  (bo_*_create with BO_FLAG_COMPILE_ONLY only needs NVRTC.)
 */
#ifndef B200OPTAS_H
#define B200OPTAS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BO_ABI_VERSION 1

typedef enum bo_status {
  BO_OK = 0,
  BO_ERR_INVALID = -1,   /* bad argument / malformed tape            */
  BO_ERR_COMPILE = -2,   /* NVRTC rejected the generated kernel       */
  BO_ERR_CUDA = -3,      /* CUDA runtime / driver error               */
  BO_ERR_NO_DEVICE = -4, /* no CUDA device (there is no CPU path)     */
  BO_ERR_UNSUPPORTED = -5/* problem too large for the tiers built     */
} bo_status;

/* Per-instance solver outcome written to status[B] by bo_solve. */
typedef enum bo_instance_status {
  BO_CONVERGED = 0,      /* scaled KKT error <= tol                                    */
  BO_ACCEPTABLE = 1,     /* stalled with KKT error <= acceptable_tol                   */
  BO_MAX_ITER = 2,       /* iteration cap reached                                      */
  BO_LINE_SEARCH = 3,    /* no acceptable step found                                   */
  BO_NUMERICAL = 4       /* NaN/Inf in function values or unfactorable KKT system      */
} bo_instance_status;

/* ---------------------------------------------------------------------------------------
 * Expression tape: the lowered form of one CasADi-style Function (what the SX virtual
 * machine interprets inside nlpsol on the reference path).  SSA-with-slot-reuse rows
 *     instr[i] = { op | (c << 8), dst, a, b }
 *   BO_OP_INPUT   work[dst] = in[b][a]            BO_OP_CONST  work[dst] = consts[a]
 *   BO_OP_OUTPUT  out[b][a] = work[dst]           BO_OP_IF_ELSE work[dst] = work[c] ? work[a] : work[b]
 *   unary/binary  work[dst] = op(work[a], work[b])
 * Opcode numbers are in bo_opcodes.h (shared with the Python front-end and the CPU oracle).
 * ------------------------------------------------------------------------------------- */
typedef struct bo_tape {
  const int32_t* instr;     /* [n_instr][4]                          */
  int64_t n_instr;
  const double* consts;     /* [n_consts]                            */
  int32_t n_consts;
  int32_t n_work;           /* work slots (after liveness reuse)     */
  int32_t n_in;
  const int32_t* in_sizes;  /* [n_in]  elements per input segment    */
  int32_t n_out;
  const int32_t* out_sizes; /* [n_out] elements per output segment   */
} bo_tape;

typedef struct bo_sparsity {  /* coordinate list of structural non-zeros */
  int32_t nnz;
  const int32_t* row;
  const int32_t* col;
} bo_sparsity;

/* Problem: min f(x,p)  s.t.  c_eq(x,p) = 0,  c_ineq(x,p) >= 0.
 * For an OpTaS `Optimization` object: c_eq = [a; h] (optimization.py:249-260,283-290),
 * c_ineq = [k; g] (optimization.py:225-247,262-281) -- NOT the +-pair stacking `v`
 * (optimization.py:27-51) the reference hands to IPOPT.
 *   fc   : (x, p)        -> f[1], c_eq[n_eq], c_ineq[n_ineq]
 *   kkt  : (x, p, y, z)  -> f[1], grad_f[nx], c_eq, c_ineq, jac_eq nz, jac_ineq nz, hess nz
 *          where hess = lower triangle of  d2/dx2 ( f - y'c_eq - z'c_ineq ).            */
typedef struct bo_problem_desc {
  int32_t nx, np, n_eq, n_ineq;
  bo_tape fc;
  bo_tape kkt;
  bo_sparsity jac_eq;    /* n_eq   x nx */
  bo_sparsity jac_ineq;  /* n_ineq x nx */
  bo_sparsity hess;      /* nx x nx, row >= col */
} bo_problem_desc;

#define BO_FLAG_COMPILE_ONLY 1u /* generate + compile (fills the cubin cache), never touch a GPU */
#define BO_FLAG_VERBOSE 2u      /* print ptxas info / cache hits to stderr                      */
#define BO_FLAG_NO_CACHE 4u     /* ignore and do not write the on-disk cubin cache              */
#define BO_FLAG_TIMING 8u       /* bracket every kernel launch with CUDA events (see *_kernel_time) */
#define BO_FLAG_PIVOTED_LDL 16u  /* factor the KKT system with Bunch-Kaufman partial pivoting (data-dependent
                                   control flow, slower) instead of the unpivoted rho-augmented LDL'  */
#define BO_FLAG_COOP 32u         /* use the cooperative tier (one instance per CTA, factor in shared memory) even for
                                   a problem small enough for the thread-per-instance sparse tier             */
#define BO_FLAG_NO_COOP 64u      /* never use the cooperative tier (large problems then run thread-per-instance) */
#define BO_FLAG_NO_QP 1024u      /* run quadratic programs through the general interior-point kernel (no dedicated QP iteration) */
#define BO_FLAG_TEAM 512u        /* use the team tier even for a problem small enough that the thread-per-instance kernel is faster */
#define BO_FLAG_PIPELINE 256u    /* host-buffer calls: cut the batch into chunks on two internal streams so that uploads, kernels
                                   and downloads overlap.  Off by default: measured slower on B200 (each chunk pays its own tail) */
#define BO_FLAG_NO_TEAM 128u     /* small dense problems: run the round-1 thread-per-instance kernel (state in thread-local
                                   memory) instead of the team tier (state in shared memory, G threads per instance) */

typedef struct bo_options {
  uint32_t flags;
  int32_t max_iter;        /* <=0: default 100                                              */
  double tol;              /* <=0: default 1e-8  (scaled KKT error, as IPOPT's `tol`)      */
  double acceptable_tol;   /* <=0: default 1e-6                                             */
  double mu_init;          /* <=0: default 0.1   (as IPOPT's `mu_init`)                    */
  double max_step;         /* cap on ||alpha*dx||_inf per iteration; 0: default = 0.5 when the tapes contain
                              sin/cos/tan (kinematics), unlimited otherwise; <0: unlimited                  */
  const char* cache_dir;   /* NULL: <directory of libb200optas.so>/_jitcache               */
  const char* include_dir; /* NULL: <directory of libb200optas.so>/csrc/jit                */
  int32_t threads_per_block; /* <=0: tier default                                           */
  int32_t max_trips;       /* <=0: default 250, or max(250, 2.5 max_iter) when max_iter is given.  Budget of solver trips per instance (one trip = at most
                              one KKT evaluation + one factorisation + one trial point); an instance
                              over budget ends with BO_MAX_ITER.  Bounds the tail latency of a batch. */
  int32_t blocks_per_sm;   /* <=0: as many as fit (occupancy); else cap on resident CTAs per SM       */
  int32_t device;          /* 0: the CUDA context current on the calling thread (none: device 0); k > 0: device
                              ordinal k-1.  The handle stays bound to that device's primary context.          */
  int32_t reserved[4];
} bo_options;

typedef struct bo_problem bo_problem;   /* opaque, owned by the library */
typedef struct bo_function bo_function; /* opaque, owned by the library */

/* Library identity. */
int bo_abi_version(void);
const char* bo_last_error(void);        /* thread-local; valid until the next call on this thread */
int bo_device_count(void);              /* number of usable CUDA devices (0 => solve/eval fail)   */

/* Replaces optas/solver.py:382 (nlpsol/qpsol factory): lowers the tapes to a fused sm_100a
 * solver kernel (NVRTC), loads it on the current device.                                    */
int bo_problem_create(const bo_problem_desc* desc, const bo_options* opts, bo_problem** out);
int bo_problem_destroy(bo_problem* prob);

/* Generated CUDA source of the problem's kernels (for inspection, offline nvcc checks and
 * the host-compiled test harness).  Returns the length; copies at most cap-1 bytes + NUL.  */
int64_t bo_problem_source(const bo_problem* prob, char* buf, int64_t cap);
/* Tables of the sparse KKT factorisation (elimination order, fill pattern, update program) built by
 * the symbolic analysis at create time; 0 entries for problems small enough for the dense path.
 * Returns the number of int32 entries; copies them if cap is large enough.                    */
int64_t bo_problem_ldl_table(const bo_problem* prob, int32_t* buf, int64_t cap);
/* Constants of the interpreted tapes (table-driven tier for very large problems); 0 entries otherwise. */
int64_t bo_problem_dtable(const bo_problem* prob, double* buf, int64_t cap);
/* Resource usage of the compiled solver kernel: regs/thread, bytes local (spill), static smem. */
int bo_problem_kernel_info(const bo_problem* prob, int32_t* regs, int32_t* local_bytes, int32_t* smem_bytes);
/* Which tier the problem was lowered to and the sizes of its plan (diagnostics; copies min(cap, BO_TIER_INFO_LEN)):
 *  [0] tier 0 dense / 1 sparse / 2 large (thread per instance)  3 cooperative (CTA per instance)  4 team (G threads
 *      in G warps per instance, state in shared memory; then [10] = G, [8] / [9] = largest / unsliced cost of the kkt
 *      slices, [15] / [16] the same for fc, [18] = shared-memory doubles per instance)  5 / 6 the dense / sparse
 *      thread-per-instance tier running the QP iteration (quadratic cost, linear constraints)
 *  [1] threads per CTA  [2] dynamic shared memory bytes  [3] elimination-tree levels  [4]/[5] sub-tapes of fc / kkt
 *  [6] parameter-only values of kkt  [7] partial-sum slots  [8] longest kkt sub-tape  [9] kkt instructions over all
 *  sub-tapes  [10]/[11] lanes per factor target / solve row  [12] multiply-adds per factorisation  [13] work slots
 *  [14] kkt components  [15] longest fc sub-tape  [16] fc instructions  [17] kkt once-per-instance instructions
 *  [18] per-CTA global workspace (doubles)  [19] factor values (doubles)  [20] resident CTAs per SM  [21] SMs
 *  [22] work slots of fc  [23] warps of the factor program  [24]/[25] steps of the longest factor / solve lane stream
 *  [26]/[27] stride of the shared-memory work arrays of kkt / fc (0: thread-local)  [28] elimination segments
 *  [29] tapes compiled per component class (1) or interpreted (0)  [30]/[31] classes / generated rows of kkt         */
#define BO_TIER_INFO_LEN 32
int bo_problem_tier_info(const bo_problem* prob, int64_t* info, int32_t cap);

/* The options in effect after defaults were resolved (max_iter, max_trips, tol, max_step ...): what the reference's
 * `solver_options` dict (optas/solver.py:333-384) became on this back-end.                                    */
int bo_problem_options(const bo_problem* prob, bo_options* out);

/* Replaces optas/solver.py:395-396 (one nlpsol call + stats) for B instances at once.
 *   p   [B][np]        parameters                      (may be NULL iff np == 0)
 *   x0  [B][nx]        initial seed, NULL = zeros      (optas/solver.py:76)
 *   x   [B][nx]   out  solution
 *   lam [B][n_eq+n_ineq] out, multipliers [y; z] of c_eq / c_ineq, or NULL
 *   f   [B]       out  cost at x, or NULL
 *   status, iters [B] out (bo_instance_status / iteration count), or NULL
 *   kkt_res [B]   out  scaled KKT error at x, or NULL                                       */
int bo_solve(bo_problem* prob, int64_t B, const double* p, const double* x0, double* x, double* lam,
             double* f, int32_t* status, int32_t* iters, double* kkt_res, void* cuda_stream);

/* Average device time (ms) of the solver-kernel launches since the last call, measured with
 * CUDA events on the launching stream; resets the counters.  n_launches may be NULL.       */
int bo_problem_kernel_time(bo_problem* prob, double* ms_total, int64_t* n_launches);

/* Batched `SXContainer.dict2vec` (optas/sx_container.py:113-123; called by Solver.reset_parameters / reset_initial_seed,
 * optas/solver.py:103-116): gathers n_seg label arrays into the row-major [B][total] matrix bo_solve takes.  Label k is
 * an m_k x n_k matrix per instance; src[k] points at element (0, 0, 0) of a [B][m_k][n_k] array whose strides in doubles are
 * stride[3k .. 3k+2] = (batch, row, column) -- any view, no contiguity needed; batch stride 0 broadcasts one matrix over the
 * batch -- or is NULL (label missing: zeros, as the reference does).  The destination
 * columns off_k .. off_k + m_k n_k - 1 receive the COLUMN-major flattening of each instance's matrix, which is the
 * reference's vec() layout.  Host memory only; n_threads <= 0: one thread per core up to 16.                          */
int bo_pack_rows(double* dst, int64_t B, int64_t total, int32_t n_seg, const double* const* src, const int64_t* stride,
                 const int32_t* m, const int32_t* n, const int64_t* off, int32_t n_threads);

/* Replaces optas/models.py:786-787 (`Function.map(n)` evaluation of an FK / Jacobian / cost
 * expression graph over n columns): a streaming kernel, one instance per thread, inputs and
 * outputs staged through shared memory with bulk async copies (TMA).
 *   in[k]  [B][in_sizes[k]]   out[k]  [B][out_sizes[k]]                                     */
int bo_function_create(const bo_tape* tape, const bo_options* opts, bo_function** out);
int bo_function_destroy(bo_function* fn);
int bo_function_eval(bo_function* fn, int64_t B, const double* const* in, double* const* out, void* cuda_stream);
int64_t bo_function_source(const bo_function* fn, char* buf, int64_t cap);
int bo_function_kernel_info(const bo_function* fn, int32_t* regs, int32_t* local_bytes, int32_t* smem_bytes);
int bo_function_kernel_time(bo_function* fn, double* ms_total, int64_t* n_launches);

#ifdef __cplusplus
}
#endif
#endif /* B200OPTAS_H */
