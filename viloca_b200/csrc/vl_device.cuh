// Device-side data model and PTX helpers of the CAVI engine (sm_100a).
//
// Vocabulary follows the reference (viloca/local_haplotype_inference): a *window* holds N
// unique reads of length L over the alphabet "ACGT-" (B = 5); a *restart* is one CAVI run on
// a window from one seeded initial state; K is the number of clusters (haplotypes).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define VL_B 5
#define VL_LOG4 1.3862943611198906   // log(B - 1)

// tile geometry (see DESIGN.md "Kernels")
#define VL_HT_POS 16     // positions per haplotype-update CTA
#define VL_H_NC 32       // reads per TMA stage of the haplotype-update kernel
#define VL_ZT_READS 64   // reads per responsibility-update CTA
#define VL_Z_PC 8        // positions per TMA stage of the responsibility-update kernel
#define VL_STAGES 3
#define VL_THREADS 128
#define VL_ZP_EXTRA 8    // scalars appended to each per-tile column-sum record
#define VL_HP_STRIDE 4

__host__ __device__ constexpr int vl_k8(int kt) { return 8 * kt; }
// Row stride (in doubles) of every K-major row.  K8 + 4 makes (4 consecutive rows) x (8
// consecutive columns) 64-bit shared-memory reads of one m8n8k4 B fragment hit 32 distinct
// banks per half-warp: 2*KS mod 32 is 8 or 24.
__host__ __device__ constexpr int vl_ks(int kt) { return 8 * kt + 4; }

// Per-restart running state: the scalar part of state_curr_dict plus the loop variables of
// run_cavi (use_quality_scores/cavi.py:114-169).
struct VlState {
  double alpha0, thr;
  double a0, b0, c0, d0;                    // state_init: gamma_a, gamma_b, theta_c, theta_d
  double lnB0, betaln_ab0, betaln_cd0;      // cavi.py:98-105
  double a, b, c, d;                        // updated gamma_a, gamma_b, theta_c, theta_d
  double mlg0, mlg1, mlt0, mlt1;            // mean_log_gamma, mean_log_theta
  double dsum_alpha, dsum_ab, dsum_cd;      // the (stale) digamma sums, cavi.py:122-128
  double elbo, t_data, t_pi, t_cluster, t_haplo, t_gamma, t_theta;
  double h_last, h_prev;                    // history_elbo[-1], history_elbo[-2]
  int iter, n_updates, n_hist, converged, status, active, max_iter, pad_;
};

// Read-only description of one (window, restart) problem.
struct VlProblem {
  int mode, N, L, K;
  int Npad, Ls, KS, KT;
  int n_ztiles, n_htiles, window, restart;
  const uint8_t* codes;    // [Npad][Ls]  0..4, 255 = N
  const double* lt;        // [Npad][Ls]  log theta            (uqs)
  const double* lo;        // [Npad][Ls]  log((1-theta)/(B-1)) (uqs)
  const double* w;         // [Npad]      read multiplicity, 0 in the padding
  const double* mcount;    // [Npad]      number of called (non-N) positions
  const uint8_t* ref;      // [Ls]
  double* z;               // [Npad][KS]       mean_cluster
  double* h;               // [Ls][B][KS]      mean_haplo, position-major
  double* zpart;           // [n_ztiles][8*KT + VL_ZP_EXTRA]
  double* hpart;           // [n_htiles][VL_HP_STRIDE]
  double* alpha;           // [8*KT]
  double* mlp;             // [8*KT]  mean_log_pi
  double* hist;            // [max_iter] history_elbo
  VlState* st;
};

// ---------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t vl_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// FP64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col).  Lane (g = lane>>2, q = lane&3)
// holds A[g][q], B[q][g], D[g][2q], D[g][2q+1].  SASS: DMMA.8x8x4.
__device__ __forceinline__ void vl_dmma(double (&d)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

__device__ __forceinline__ void vl_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(vl_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void vl_fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void vl_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
               :: "r"(vl_smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (1-D), completion counted in bytes on an mbarrier.
// SASS: UBLKCP.  dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void vl_tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
      :: "r"(vl_smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(vl_smem_u32(bar))
      : "memory");
}
// Bounded wait: a lost completion traps (CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void vl_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = vl_smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (spin > (1u << 22)) __trap();
  }
}

// digamma for x > 0: upward recurrence to x >= 10, then the asymptotic series
// (scipy.special.digamma at update_eqs.py:45,51-52 and cavi.py:123-126).
__device__ __forceinline__ double vl_digamma(double x) {
  double r = 0.0;
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double inv = 1.0 / x, i2 = inv * inv;
  const double s = i2 * (1.0 / 12 - i2 * (1.0 / 120 - i2 * (1.0 / 252 - i2 * (1.0 / 240 -
                   i2 * (1.0 / 132 - i2 * (691.0 / 32760 - i2 * (1.0 / 12)))))));
  return r + (log(x) - 0.5 * inv - s);
}

__device__ __forceinline__ double vl_betaln(double a, double b) {
  return lgamma(a) + lgamma(b) - lgamma(a + b);
}
