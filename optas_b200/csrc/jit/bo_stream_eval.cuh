// bo_stream_eval.cuh -- streaming evaluation of one expression tape over a batch (kernel "K1":
// forward kinematics / Jacobian / cost / constraint graphs), one instance per thread.
//
// Replaces `cs.Function(...).map(n)` evaluation on the reference path (optas/models.py:786-787):
// CasADi walks the n columns serially on one core; here the batch is streamed through the SMs.
// The arithmetic is FP64 straight-line code generated from the tape (bo_tape_call); this file is
// the memory system around it, written for HBM3e:
//   * global layout is row-major [B][size] per segment, so one tile of BO_TPB instances is ONE
//     contiguous byte range per segment -> moved with 1-D bulk async copies (TMA engine,
//     cp.async.bulk, SASS UBLKCP) global->shared on an mbarrier, and shared->global as bulk groups;
//     every byte crosses HBM exactly once, fully coalesced, no per-thread strided global access;
//   * 2-stage pipeline: the loads of tile k+1 are in flight while tile k computes, the stores of
//     tile k-1 drain meanwhile; persistent CTAs, grid = multiple of the SM count;
//   * threads read their operands from shared memory at stride `size` doubles (odd sizes such as
//     3, 7, 21 are bank-conflict-free).
// The generated prelude defines BO_NIN, BO_NOUT, BO_TPB, BO_IN_TOTAL, BO_OUT_TOTAL, BO_IN_SIZE[],
// BO_OUT_SIZE[] and bo_tape_call(si, so, t).
#pragma once
#include "bo_common.cuh"

struct bo_eval_args {
  const double* in[BO_NIN];
  double* out[BO_NOUT];
  int use_bulk;  // 0: all pointers are only 8-byte aligned -> plain coalesced copies
};

#ifndef BO_OUT_STAGES
#define BO_OUT_STAGES 2  // 2: outputs double-buffered like the inputs; 1: one output buffer (more resident CTAs per SM)
#endif
#define BO_IN_STAGE_DOUBLES (BO_TPB * BO_IN_TOTAL)
#define BO_OUT_STAGE_DOUBLES (BO_TPB * BO_OUT_TOTAL)

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
__device__ __forceinline__ void bo_bulk_s2g(void* dst_gmem, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(bo_smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bo_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bo_bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bo_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bo_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifdef BO_MIN_BLOCKS  // experiment knob (-DBO_MIN_BLOCKS=n via B200OPTAS_JIT_DEFINES): register cap for n resident CTAs per SM
#define BO_EVAL_BOUNDS __launch_bounds__(BO_TPB, BO_MIN_BLOCKS)
#else
#define BO_EVAL_BOUNDS __launch_bounds__(BO_TPB)
#endif
extern "C" __global__ void BO_EVAL_BOUNDS bo_eval_kernel(long long B, const bo_eval_args args) {
  extern __shared__ __align__(128) unsigned char bo_smem_raw[];
  double* const stage_base = reinterpret_cast<double*>(bo_smem_raw);
  double* const out_base = stage_base + 2 * BO_IN_STAGE_DOUBLES;
  unsigned long long* const full = reinterpret_cast<unsigned long long*>(out_base + BO_OUT_STAGES * BO_OUT_STAGE_DOUBLES);

  // shared-memory segment pointers of one stage
  auto stage_ptrs = [&](int s, double** si, double** so) {
    double* ptr = stage_base + s * BO_IN_STAGE_DOUBLES;
    BO_UNROLL
    for (int k = 0; k < BO_NIN; ++k) {
      si[k] = ptr;
      ptr += BO_TPB * BO_IN_SIZE[k];
    }
    ptr = out_base + (BO_OUT_STAGES == 2 ? s : 0) * BO_OUT_STAGE_DOUBLES;
    BO_UNROLL
    for (int k = 0; k < BO_NOUT; ++k) {
      so[k] = ptr;
      ptr += BO_TPB * BO_OUT_SIZE[k];
    }
  };

  const int tid = threadIdx.x;
  const long long n_tiles = (B + BO_TPB - 1) / BO_TPB;
  const long long n_full_tiles = B / BO_TPB;  // tiles that can use bulk copies
  const bool bulk = args.use_bulk != 0;

  if (tid == 0) {
    bo_mbar_init(&full[0], 1);
    bo_mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue_loads = [&](long long tile, int s) {  // thread 0 only; tile is a full tile
    double *si[BO_NIN], *so[BO_NOUT];
    stage_ptrs(s, si, so);
    bo_mbar_expect_tx(&full[s], (unsigned)(BO_TPB * BO_IN_TOTAL * sizeof(double)));
    BO_UNROLL
    for (int k = 0; k < BO_NIN; ++k)
      if (BO_IN_SIZE[k] > 0)
        bo_bulk_g2s(si[k], args.in[k] + tile * BO_TPB * BO_IN_SIZE[k], (unsigned)(BO_TPB * BO_IN_SIZE[k] * sizeof(double)),
                  &full[s]);
  };

  long long it = 0;  // tiles processed by this CTA so far (selects stage and mbarrier parity)
  if (bulk && tid == 0 && (long long)blockIdx.x < n_full_tiles) issue_loads(blockIdx.x, 0);

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int s = (int)(it & 1);
    double *si[BO_NIN], *so[BO_NOUT];
    stage_ptrs(s, si, so);
    const long long next = tile + gridDim.x;
    const bool tile_bulk = bulk && tile < n_full_tiles;
    const int count = (int)((B - tile * BO_TPB) < (long long)BO_TPB ? (B - tile * BO_TPB) : (long long)BO_TPB);

    if (tid == 0) {
      // stores of the tile that last used this output buffer (two tiles ago, or the previous tile when there is
      // only one output buffer) must have finished reading smem
      bo_bulk_wait_read<BO_OUT_STAGES - 1>();
    }
    __syncthreads();  // everyone is done computing the previous tile (stage s^1 inputs are free)
    if (bulk && tid == 0 && next < n_full_tiles) issue_loads(next, s ^ 1);

    if (tile_bulk) {
      bo_mbar_wait(&full[s], (unsigned)((it >> 1) & 1));
    } else {
      BO_UNROLL
      for (int k = 0; k < BO_NIN; ++k) {
        const double* src = args.in[k] + tile * BO_TPB * BO_IN_SIZE[k];
        for (int e = tid; e < count * BO_IN_SIZE[k]; e += BO_TPB) si[k][e] = src[e];
      }
      __syncthreads();
    }

    if (tid < count) bo_tape_call(si, so, tid);

    if (tile_bulk) {
      bo_fence_async_smem();  // make generic-proxy smem writes visible to the bulk-copy engine
      __syncthreads();
      if (tid == 0) {
        BO_UNROLL
        for (int k = 0; k < BO_NOUT; ++k)
          if (BO_OUT_SIZE[k] > 0)
            bo_bulk_s2g(args.out[k] + tile * BO_TPB * BO_OUT_SIZE[k], so[k],
                      (unsigned)(BO_TPB * BO_OUT_SIZE[k] * sizeof(double)));
        bo_bulk_commit();
      }
    } else {
      __syncthreads();
      BO_UNROLL
      for (int k = 0; k < BO_NOUT; ++k) {
        double* dst = args.out[k] + tile * BO_TPB * BO_OUT_SIZE[k];
        for (int e = tid; e < count * BO_OUT_SIZE[k]; e += BO_TPB) dst[e] = so[k][e];
      }
      if (tid == 0) bo_bulk_commit();  // empty group keeps the wait_group.read<1> bookkeeping uniform
    }
  }
  if (tid == 0) bo_bulk_wait_all();
}
