// FP64 ceiling microbenchmark for B200 (sm_100a).
//
// MEASURED_PEAKS.json carries no FP64 entry, so the roofline denominator for the
// two CAVI contractions is measured here: (1) a register-resident DMMA m8n8k4
// loop (the instruction the contraction kernels issue), (2) a register-resident
// DFMA loop (the vector pipe), both timed with CUDA events.  A cuBLAS DGEMM
// ceiling is measured separately from Python (tools/dgemm_peak.py).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void __launch_bounds__(1024) dmma_loop(double* out, int iters, double a0, double b0) {
  double acc[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma884(acc[i][0], acc[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(1024) dfma_loop(double* out, int iters, double a0, double b0) {
  double acc[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"results\": [\n", p.name, sms);
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
  const int iters = 4096;
  bool first = true;
  // DMMA: sweep warps per SM (block size) x CTAs/SM with 8 independent accumulators
  int cfgs[][2] = {{128, 1}, {256, 1}, {512, 1}, {1024, 1}, {128, 2}, {256, 2}, {128, 4}};
  for (auto& c : cfgs) {
    int threads = c[0], cps = c[1];
    float ms = time_ms([&] { dmma_loop<8><<<sms * cps, threads>>>(out, iters, 1.0, 1e-3); }, 5);
    double flops = 2.0 * 8 * 8 * 4 * 8.0 * iters * (threads / 32) * sms * cps;
    printf("%s{\"kind\": \"dmma884\", \"threads\": %d, \"ctas_per_sm\": %d, \"nacc\": 8, \"ms\": %.4f, \"tflops\": %.3f}",
           first ? "" : ",\n", threads, cps, ms, flops / ms * 1e-9);
    first = false;
  }
  {
    float ms = time_ms([&] { dmma_loop<2><<<sms, 256>>>(out, iters, 1.0, 1e-3); }, 5);
    double flops = 2.0 * 256 * 2.0 * iters * 8 * sms;
    printf(",\n{\"kind\": \"dmma884\", \"threads\": 256, \"ctas_per_sm\": 1, \"nacc\": 2, \"ms\": %.4f, \"tflops\": %.3f}", ms, flops / ms * 1e-9);
    ms = time_ms([&] { dmma_loop<1><<<sms, 128>>>(out, iters, 1.0, 1e-3); }, 5);
    flops = 2.0 * 256 * 1.0 * iters * 4 * sms;
    printf(",\n{\"kind\": \"dmma884_latency\", \"threads\": 128, \"ctas_per_sm\": 1, \"nacc\": 1, \"ms\": %.4f, \"tflops\": %.3f, \"ns_per_dependent_mma\": %.3f}",
           ms, flops / ms * 1e-9, ms * 1e6 / iters);
  }
  int fcfgs[][2] = {{256, 1}, {512, 1}, {1024, 1}, {1024, 2}};
  for (auto& c : fcfgs) {
    int threads = c[0], cps = c[1];
    float ms = time_ms([&] { dfma_loop<8><<<sms * cps, threads>>>(out, iters, 0.999999, 1e-3); }, 5);
    double flops = 2.0 * 8.0 * iters * threads * sms * cps;
    printf(",\n{\"kind\": \"dfma\", \"threads\": %d, \"ctas_per_sm\": %d, \"nacc\": 8, \"ms\": %.4f, \"tflops\": %.3f}",
           threads, cps, ms, flops / ms * 1e-9);
  }
  printf("\n]}\n");
  return 0;
}
