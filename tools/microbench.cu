// Latency / issue-cost microbenchmarks behind the kernel design (sm_100a, one CTA unless stated):
//   dependent and independent DMMA m8n8k4 chains, FP64 FMA chain, exp/log/rcp chains,
//   cp.async.bulk issue cost and completion latency (L2-resident source), mbarrier try_wait,
//   L2 round trips (volatile load, atomicAdd), __threadfence, __nanosleep, __syncthreads.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/microbench tools/microbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               :: "r"(smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

#define NREP 256
__global__ void k_alu(long long* out, double* sink, double seed) {
  const int lane = threadIdx.x & 31;
  long long t0, t1;
  int slot = 0;
  double a = seed + lane * 1e-3, b = 1.0 + 1e-9 * lane;
  {  // one dependent DMMA chain
    double d[2] = {0.0, 0.0};
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < NREP; ++i) dmma(d, a, b);
    t1 = clock64();
    sink[0] = d[0] + d[1];
    if (threadIdx.x == 0) out[slot] = t1 - t0;
    ++slot;
  }
  {  // 2, 4, 8 independent chains per warp
    double d[8][2] = {};
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < NREP; ++i) { dmma(d[0], a, b); dmma(d[1], a, b); }
    t1 = clock64();
    if (threadIdx.x == 0) out[slot] = t1 - t0;
    ++slot;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < NREP; ++i) { dmma(d[0], a, b); dmma(d[1], a, b); dmma(d[2], a, b); dmma(d[3], a, b); }
    t1 = clock64();
    if (threadIdx.x == 0) out[slot] = t1 - t0;
    ++slot;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < NREP; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) dmma(d[j], a, b);
    }
    t1 = clock64();
    if (threadIdx.x == 0) out[slot] = t1 - t0;
    ++slot;
    double s = 0.0;
    for (int j = 0; j < 8; ++j) s += d[j][0] + d[j][1];
    sink[1] = s;
  }
  {  // dependent FMA chain
    double x = a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < NREP; ++i) x = fma(x, b, 1e-12);
    t1 = clock64();
    sink[2] = x;
    if (threadIdx.x == 0) out[slot] = t1 - t0;
    ++slot;
  }
  {  // dependent exp / log / reciprocal chains (CUDA library versions), 32 each
    double x = -0.5 - 1e-3 * lane;
    t0 = clock64();
    for (int i = 0; i < 32; ++i) x = -exp(x);
    t1 = clock64();
    sink[3] = x;
    if (threadIdx.x == 0) out[slot] = t1 - t0;
    ++slot;
    x = 2.5 + lane;
    t0 = clock64();
    for (int i = 0; i < 32; ++i) x = log(x) + 3.0;
    t1 = clock64();
    sink[4] = x;
    if (threadIdx.x == 0) out[slot] = t1 - t0;
    ++slot;
    x = 1.5 + lane;
    t0 = clock64();
    for (int i = 0; i < 32; ++i) x = 1.0 / x + 1.0;
    t1 = clock64();
    sink[5] = x;
    if (threadIdx.x == 0) out[slot] = t1 - t0;
    ++slot;
  }
  {  // __syncthreads x 64 (all warps of the CTA)
    __syncthreads();
    t0 = clock64();
    for (int i = 0; i < 64; ++i) __syncthreads();
    t1 = clock64();
    if (threadIdx.x == 0) out[slot] = t1 - t0;
    ++slot;
  }
}

// all four warps issue DMMAs at once: per-SM pipe throughput (4 independent chains per warp)
__global__ void k_dmma_sm(long long* out, double* sink, double seed) {
  const int lane = threadIdx.x & 31;
  double a = seed + lane * 1e-3, b = 1.0 + 1e-9 * lane;
  double d[4][2] = {};
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < NREP; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) dmma(d[j], a, b);
  }
  __syncthreads();
  const long long t1 = clock64();
  double s = 0.0;
  for (int j = 0; j < 4; ++j) s += d[j][0] + d[j][1];
  sink[threadIdx.x] = s;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

// memory-system round trips by thread 0
__global__ void k_mem(long long* out, int* flags, const double* src, int n_copy_bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bars[16];
  long long t0, t1;
  int slot = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 16; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile int* vf = flags;
    int acc = 0;
    t0 = clock64();
    for (int i = 0; i < 64; ++i) acc += vf[(acc & 1) + 32 * (i & 7)];      // dependent volatile loads (L2)
    t1 = clock64();
    out[slot++] = t1 - t0; flags[300] = acc;
    t0 = clock64();
    for (int i = 0; i < 64; ++i) acc += atomicAdd(&flags[64 + (acc & 1)], 1);   // dependent atomics
    t1 = clock64();
    out[slot++] = t1 - t0; flags[301] = acc;
    t0 = clock64();
    for (int i = 0; i < 64; ++i) { flags[128 + i] = i; __threadfence(); }
    t1 = clock64();
    out[slot++] = t1 - t0;
    t0 = clock64();
    for (int i = 0; i < 64; ++i) __nanosleep(40);
    t1 = clock64();
    out[slot++] = t1 - t0;
    // one bulk copy: issue cost and completion latency
    t0 = clock64();
    mbar_expect_tx(&bars[0], n_copy_bytes);
    tma_load_1d(smem, src, n_copy_bytes, &bars[0]);
    t1 = clock64();
    mbar_wait(&bars[0], 0);
    const long long t2 = clock64();
    out[slot++] = t1 - t0;
    out[slot++] = t2 - t0;
    // 8 bulk copies back to back on 8 barriers, then wait for all
    t0 = clock64();
    for (int i = 0; i < 8; ++i) {
      mbar_expect_tx(&bars[1 + i], n_copy_bytes);
      tma_load_1d(smem + (size_t)(i + 1) * n_copy_bytes, src + (size_t)(i + 1) * n_copy_bytes / 8, n_copy_bytes, &bars[1 + i]);
    }
    t1 = clock64();
    for (int i = 0; i < 8; ++i) mbar_wait(&bars[1 + i], 0);
    const long long t3 = clock64();
    out[slot++] = t1 - t0;
    out[slot++] = t3 - t0;
    // try_wait on an already completed barrier
    t0 = clock64();
    for (int i = 0; i < 64; ++i) mbar_wait(&bars[0], 0);
    t1 = clock64();
    out[slot++] = t1 - t0;
  }
}

// flag ping-pong between two CTAs (on different SMs): one-way signalling latency through L2
__global__ void k_pingpong(long long* out, int* flags) {
  volatile int* f = flags;
  if (threadIdx.x != 0) return;
  const int me = blockIdx.x;
  const long long t0 = clock64();
  for (int i = 1; i <= 200; ++i) {
    if (me == 0) { f[0] = i; __threadfence(); while (f[32] < i) { } }
    else { while (f[0] < i) { } f[32] = i; __threadfence(); }
  }
  const long long t1 = clock64();
  if (me == 0) out[0] = t1 - t0;
}

int main() {
  long long* out; double* sink; int* flags; double* src;
  cudaMallocManaged(&out, 64 * sizeof(long long));
  cudaMalloc(&sink, 1024 * sizeof(double));
  cudaMalloc(&flags, 4096 * sizeof(int));
  cudaMemset(flags, 0, 4096 * sizeof(int));
  cudaMalloc(&src, 1 << 20);
  cudaMemset(src, 0, 1 << 20);
  for (int rep = 0; rep < 2; ++rep) {
    k_alu<<<1, 128>>>(out, sink, 0.5);
    cudaDeviceSynchronize();
  }
  const char* alu[] = {"DMMA dependent chain", "DMMA 2 chains", "DMMA 4 chains", "DMMA 8 chains", "DFMA dependent chain",
                       "exp chain (32)", "log chain (32)", "rcp chain (32)", "__syncthreads (64)"};
  const int alun[] = {NREP, 2 * NREP, 4 * NREP, 8 * NREP, NREP, 32, 32, 32, 64};
  for (int i = 0; i < 9; ++i) printf("%-28s %8lld cycles  %8.1f per op\n", alu[i], out[i], (double)out[i] / alun[i]);
  for (int rep = 0; rep < 2; ++rep) { k_dmma_sm<<<1, 128>>>(out, sink, 0.5); cudaDeviceSynchronize(); }
  printf("%-28s %8lld cycles  %8.1f per DMMA per warp (4 warps x 4 chains)\n", "DMMA 4 warps", out[0], (double)out[0] / (4 * NREP));
  for (int rep = 0; rep < 2; ++rep) { k_dmma_sm<<<1, 256>>>(out, sink, 0.5); cudaDeviceSynchronize(); }
  printf("%-28s %8lld cycles  %8.1f per DMMA per warp (8 warps x 4 chains)\n", "DMMA 8 warps", out[0], (double)out[0] / (4 * NREP));
  for (int bytes : {2048, 8192}) {
    cudaFuncSetAttribute(k_mem, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) { k_mem<<<1, 128, 9 * 8192 + 128>>>(out, flags, src, bytes); cudaDeviceSynchronize(); }
    printf("-- bulk copy of %d bytes --\n", bytes);
    const char* mem[] = {"volatile load (64 dep)", "atomicAdd (64 dep)", "store+threadfence (64)", "nanosleep(40) (64)",
                         "bulk copy issue", "bulk copy issue->complete", "8 bulk copies issue", "8 bulk copies ->complete",
                         "try_wait completed (64)"};
    const int memn[] = {64, 64, 64, 64, 1, 1, 8, 1, 64};
    for (int i = 0; i < 9; ++i) printf("%-28s %8lld cycles  %8.1f per op\n", mem[i], out[i], (double)out[i] / memn[i]);
  }
  cudaMemset(flags, 0, 4096 * sizeof(int));
  k_pingpong<<<2, 32>>>(out, flags);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-28s %8lld cycles  %8.1f per round trip (2 one-way signals)   [%s]\n", "flag ping-pong (200)", out[0], (double)out[0] / 200,
         cudaGetErrorString(e));
  return 0;
}
