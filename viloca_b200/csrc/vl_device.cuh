// Device-side data model and PTX helpers of the CAVI engine (sm_100a).
//
// Vocabulary follows the reference (viloca/local_haplotype_inference): a *window* holds N
// unique reads of length L over the alphabet "ACGT-" (B = 5); a *restart* is one CAVI run on
// a window from one seeded initial state; K is the number of clusters (haplotypes).
//
// Formulation.  The reference's operand E[n,l,b] (preparation.py:123-157) has
// two distinct values per (read, position): log theta at the called base c and
// log((1-theta)/(B-1)) elsewhere.  With d = log theta - log((1-theta)/(B-1)) and
// g[k,l,b] = 1 - h[k,l,b] = sum_{b' != b} h[k,l,b']:
//   contraction #1   sum_{l,b} E h        = ltsum[n] - Y[n,k],  Y = sum_l d[n,l] g[k,l,c[n,l]]
//   contraction #2   sum_n w z E [b]      = (b-independent) + sum_n w z d [c[n,l] == b]
// The b-independent part cancels in the softmax over bases.  Both are contractions with the
// one-hot-expanded matrix D[n,(l,b)] = d[n,l] [c[n,l] == b], which has one non-zero per called
// (read, position).  Its *popular* columns -- the most frequent base of every position, plus
// every other (position, base) pair that enough reads carry -- form the dense operand
// Dm[n,j] that is contracted on the FP64 tensor pipe (about L columns instead of the 5 L of E);
// the few remaining entries (sequencing errors, rare variants) are kept in two sparse lists
// (by read, by (position, base)).  Clusters whose responsibilities underflowed to exactly 0
// for every read are removed from the dense work: the *layout* lists the live clusters plus
// one ghost slot that carries the (identical) values of all dead clusters with their
// multiplicity.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define VL_B 5
#define VL_LOG4 1.3862943611198906   // log(B - 1)

// tile geometry (see DESIGN.md "Kernels")
#define VL_PT 16         // columns per operand tile (= columns per haplotype-update CTA)
#define VL_ZT_READS 64   // reads per responsibility-update CTA
#define VL_H_NC 32       // reads per stage of the haplotype-update kernel (8 per warp)
#define VL_STAGES 3      // pipeline stages the shared memory of a kernel class is sized for (widest layout)
#ifndef VL_ZSTAGES
#define VL_ZSTAGES 3     // the same for the responsibility-update kernel
#endif
#ifndef VL_MAX_STAGES
#define VL_MAX_STAGES 24 // most stages in flight (narrow layouts, or few restarts left: see vl_launch_chunk)
#endif
#define VL_THREADS 128
#define VL_MAX_KT 16     // 8-slot tiles of a layout (VL_MAX_CLUSTERS / 8)
#define VL_HSPLIT_READS 512   // a haplotype tile's reads are shared by Npad / 512 CTAs ...
#define VL_MAX_HSPLIT 8       // ... at most 8 (windows of fewer than 1024 padded reads: one CTA, as before)
#define VL_FINE_ZT_MAX_READS 16384   // windows with split haplotype tiles and at most this many padded reads use read tiles of 32 reads
#define VL_ZP_EXTRA 8    // scalars appended to each per-tile column-sum record: data, ent, c, d
#define VL_HP_STRIDE 4

// Row stride (in doubles) of every slot-major row of a layout with kt tiles.  8*kt + 4 makes
// (4 consecutive rows) x (8 consecutive columns) 64-bit shared-memory reads of one m8n8k4 B
// fragment hit 16 distinct 8-byte banks per half-warp.
__host__ __device__ constexpr int vl_ks(int kt) { return 8 * kt + 4; }

// Dense operand tiles: element (read n, column j) of Dm lives at
//   ((j / 16) * Npad + n) * 16 + ((j % 16 + 4 * (n % 4)) % 16)
// i.e. column-tile-major, one 128-byte line per read, rotated by 4 doubles per read so that
// both fragment shapes (8 reads x 4 columns, 4 reads x 8 columns) read 16 distinct banks.
// All columns of one position lie in the same tile.
__host__ __device__ __forceinline__ int vl_dm_col(int n, int p) { return (p + 4 * (n & 3)) & 15; }

// Per-restart running state: the scalar part of state_curr_dict plus the loop variables of
// run_cavi (use_quality_scores/cavi.py:114-169) and the live-cluster layout.
struct VlState {
  double alpha0, thr;
  double a0, b0, c0, d0;                    // state_init: gamma_a, gamma_b, theta_c, theta_d
  double lnB0, betaln_ab0, betaln_cd0;      // cavi.py:98-105
  double a, b, c, d;                        // updated gamma_a, gamma_b, theta_c, theta_d
  double mlg0, mlg1, mlt0, mlt1;            // mean_log_gamma, mean_log_theta
  double dsum_alpha, dsum_ab, dsum_cd;      // the (stale) digamma sums, cavi.py:122-128
  double elbo, t_data, t_pi, t_cluster, t_haplo, t_gamma, t_theta;
  double h_last, h_prev;                    // history_elbo[-1], history_elbo[-2]
  double ghost_mult;                        // number of dead clusters the ghost slot stands for
  int iter, n_updates, n_hist, converged, status, active, max_iter;
  int ktl;        // tiles of the current layout (h, g, gm and the next z are stored in it)
  int ktl_z;      // tiles of the layout z is stored in (gat maps current slots to its slots)
  int nlay;       // valid slots of the current layout: live clusters (+ 1 ghost)
  int ghost;      // slot of the ghost, or -1 when every cluster is live
  int nlay_z, ghost_z;   // the same for the layout z (and the last h) are stored in
  int skip;       // 1 = remove dead clusters from the layout (exact), 0 = always dense
  int stalled;    // layout grew past the kernel class this restart is scheduled in
  int ticket;     // read tiles of the current responsibility update that have finished
  int pub;        // updates completed AND visible: written last by the scalar update (fused chunks poll it)
  int ktring[2];  // ktring[u & 1] = tiles of the layout update u stores z / w z in (set by the update before it)
  int pad_;
};

// Scheduling word of one (window, restart) problem inside a chunk launch (vl_sched_kernel); zeroed by the
// host before every chunk.  Phase 2k = haplotype tiles of the chunk's update k, 2k + 1 = its read tiles.
struct VlCtl {
  unsigned long long claim;   // (phase << 32) | next tile of that phase to hand out; VL_PHASE_LEFT once the restart left the chunk
  int done;                   // tiles of the current phase that have finished
  int pad_;
};
#define VL_PHASE_LEFT 0x7fffffffu

// Read-only description of one (window, restart) problem.
struct VlProblem {
  int mode, N, L, K;
  int Npad, Ls, KT, NCs;   // NCs: dense columns incl. padding, a multiple of 16
  int n_ztiles, n_htiles, window, restart;
  // window tensors (shared by the restarts of a window)
  const double* dm;        // [NCs/16][Npad][16] dense operand, see vl_dm_col
  const double* w;         // [Npad]  read multiplicity, 0 in the padding
  const double* ltsum;     // [Npad]  uqs: sum over called positions of log theta
  const double* mcount;    // [Npad]  number of called (non-N) positions
  const uint8_t* ref;      // [Ls]
  const int* col_lb;       // [NCs]   dense column -> l * 5 + b, or -1 (padding)
  const int* poscol;       // [Ls*5]  l * 5 + b -> dense column, or -1 (pair kept in the sparse lists)
  const int* tile_pos;     // [NCs]   the positions whose columns lie in tile j / 16, then -1
  const int* sr_ptr;       // [Npad+1] sparse entries by read ...
  const int* sr_idx;       //   l * 5 + c
  const double* sr_d;      //   d[n,l]
  const int* sp_ptr;       // [Ls*5+1] ... and by (position, base), reads ascending
  const int* sp_n;         //   n
  const double* sp_d;      //   d[n,l]
  // restart tensors; "slot" = column of the current layout
  double* z;               // [Npad][vl_ks(ktl_z)]   mean_cluster
  double* wz;              // [Npad][vl_ks(ktl_z)]   w_n * mean_cluster (operand of the haplotype update)
  double* h;               // [Ls][5][vl_ks(ktl)]    mean_haplo
  double* g;               // [Ls][5][vl_ks(ktl)]    1 - mean_haplo, summed from the other bases
  double* gm;              // [NCs][vl_ks(ktl)]      g at the (position, base) of each dense column
  double* zpart;           // [n_ztiles][8*KT + VL_ZP_EXTRA]
  double* hpart;           // [n_htiles][VL_HP_STRIDE]
  double* alpha;           // [8*KT]  per cluster
  double* mlp;             // [8*KT]  mean_log_pi per cluster
  double* mlp_lay;         // [8*KT]  mean_log_pi per slot
  int* lay;                // [8*KT]  slot -> cluster (ghost: one of the dead clusters)
  int* gat;                // [8*KT]  slot -> slot of the layout z is stored in
  int* lay_z;              // [8*KT]  slot -> cluster for the layout z is stored in
  double* hist;            // [max_iter] history_elbo
  VlState* st;
  // large windows only (kept behind the fields every tile reads, whose cache lines stay as they were)
  double* tsplit;          // [n_htiles][n_hsplit][16 * 8*KT]  partial sums over reads of a split tile (n_hsplit > 1)
  int* hsplit_count;       // [n_htiles]  arrivals at a split tile in the current update
  int n_hsplit, zt_reads;  // CTAs that share the reads of one haplotype tile (1 unless the window is large); reads per read tile (64 or 32)
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
// An mbarrier must be invalidated before its memory is initialised again (persistent CTAs run tile after tile).
__device__ __forceinline__ void vl_mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];\n" :: "r"(vl_smem_u32(bar)) : "memory");
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
// (scipy.special.digamma at update_eqs.py:45,51-52 and cavi.py:123-126).  The (at most ten)
// reciprocals of the recurrence are independent of each other, so they are issued together
// and only then subtracted in order: same value as the sequential loop, a tenth of its latency.
__device__ __forceinline__ double vl_digamma(double x) {
  double r = 0.0;
  if (x < 10.0) {
    double inv[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const bool on = x < 10.0;
      inv[i] = on ? 1.0 / x : 0.0;
      x = on ? x + 1.0 : x;
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) r -= inv[i];
  }
  const double inv = 1.0 / x, i2 = inv * inv;
  const double s = i2 * (1.0 / 12 - i2 * (1.0 / 120 - i2 * (1.0 / 252 - i2 * (1.0 / 240 -
                   i2 * (1.0 / 132 - i2 * (691.0 / 32760 - i2 * (1.0 / 12)))))));
  return r + (log(x) - 0.5 * inv - s);
}

// digamma and log-gamma of x > 0 together (scipy.special.digamma / gammaln / betaln of
// update_eqs.py:40-53, elbo_eqs.py:56-62,81-88): both shift x upward to y >= 10 and use their
// asymptotic series there; they share the reciprocals of the shift and log(y), and the shift of
// lgamma is one log of the product (x)(x+1)...(y-1) <= 3.4e11.  FP64 latency bounds the scalar
// update of an iteration (a handful of warps, serial chains), and CUDA's lgamma() alone is a
// several-thousand-cycle chain; this pair costs two overlapping logs and a short series.
// Absolute error about 1e-15 (truncation of the Stirling series at y^-13: 3e-17).
__device__ __forceinline__ void vl_psi_lgamma(double x, double& psi, double& lg) {
  double r = 0.0, prod = 1.0;
  if (x < 10.0) {
    double inv[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const bool on = x < 10.0;
      inv[i] = on ? 1.0 / x : 0.0;
      prod = on ? prod * x : prod;
      x = on ? x + 1.0 : x;
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) r -= inv[i];
  }
  const double ly = log(x), lp = log(prod);
  const double inv = 1.0 / x, i2 = inv * inv;
  const double sp = i2 * (1.0 / 12 - i2 * (1.0 / 120 - i2 * (1.0 / 252 - i2 * (1.0 / 240 -
                    i2 * (1.0 / 132 - i2 * (691.0 / 32760 - i2 * (1.0 / 12)))))));
  psi = r + (ly - 0.5 * inv - sp);
  const double sl = inv * (1.0 / 12 - i2 * (1.0 / 360 - i2 * (1.0 / 1260 - i2 * (1.0 / 1680 -
                    i2 * (1.0 / 1188 - i2 * (691.0 / 360360 - i2 * (1.0 / 156)))))));
  lg = ((x - 0.5) * ly - x + 0.91893853320467274178 + sl) - lp;
}

__device__ __forceinline__ double vl_lgamma(double x) {
  double psi, lg;
  vl_psi_lgamma(x, psi, lg);
  return lg;
}

__device__ __forceinline__ double vl_betaln(double a, double b) {
  return vl_lgamma(a) + vl_lgamma(b) - vl_lgamma(a + b);
}
