// fp64_microbench.cu -- what the FP64 pipes of one B200 deliver, measured two ways (BASELINE.md section 5 asks for both
// before any compute fraction is quoted):
//   1. scalar DFMA: 8 independent chains per thread (enough ILP to cover the pipe latency), all SMs, persistent;
//   2. FP64 tensor path: mma.sync.aligned.m8n8k4.row.col.f64 (the only FP64 MMA on this architecture; tcgen05 has no f64
//      kind), 8 independent accumulator tiles per warp.
// Prints TFLOP/s for both.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/fp64_microbench.cu -o tools/_build/fp64_microbench
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int CHAINS = 8;

__global__ void __launch_bounds__(256) fma_kernel(double* out, int iters, double a, double b) {
  double v[CHAINS];
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) v[k] = threadIdx.x * 1e-3 + k;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) v[k] = fma(v[k], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) s += v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dmma_kernel(double* out, int iters, double a0, double b0) {
  double c[CHAINS][2];
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) c[k][0] = c[k][1] = 0.0;
  const double a = a0 + threadIdx.x * 1e-6, b = b0;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < CHAINS; ++k)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < CHAINS; ++k) s += c[k][0] + c[k][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int dev = 0, sms = 0;
  CHECK(cudaGetDevice(&dev));
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int threads = 256, blocks = sms * 8, iters = 1 << 15;
  double* out;
  CHECK(cudaMalloc(&out, sizeof(double) * threads * blocks));
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  float ms;
  for (int pass = 0; pass < 2; ++pass) {  // first pass warms up
    CHECK(cudaEventRecord(e0));
    fma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    if (pass) printf("{\"kernel\": \"dfma\", \"tflops\": %.3f, \"ms\": %.3f, \"sms\": %d, \"chains_per_thread\": %d}\n",
                     2.0 * CHAINS * (double)iters * threads * blocks / (ms * 1e-3) / 1e12, ms, sms, CHAINS);
  }
  for (int pass = 0; pass < 2; ++pass) {
    CHECK(cudaEventRecord(e0));
    dmma_kernel<<<blocks, threads>>>(out, iters, 0.5, 0.25);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaEventSynchronize(e1));
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    // one m8n8k4 = 8 * 8 * 4 multiply-adds = 512 flop per warp
    if (pass) printf("{\"kernel\": \"dmma m8n8k4\", \"tflops\": %.3f, \"ms\": %.3f, \"mma_per_warp\": %d}\n",
                     512.0 * CHAINS * (double)iters * (threads / 32) * blocks / (ms * 1e-3) / 1e12, ms, CHAINS * iters);
  }
  CHECK(cudaFree(out));
  return 0;
}
