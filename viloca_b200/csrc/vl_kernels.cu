// CAVI kernels for sm_100a: the two dense FP64 contractions of one CAVI iteration on the
// FP64 tensor pipe (DMMA m8n8k4) with the softmax / ELBO reductions fused into their
// epilogues, plus the per-restart scalar update that closes the iteration.
//
// One iteration of a (window, restart) problem = three launches over the whole batch:
//   vl_haplo_kernel    update_mean_haplo   (uqs update_eqs.py:73-101, lep update_eqs.py:115-159)
//                      + the sums update_a_and_b / elbo_haplo need (update_eqs.py:103-112,
//                      elbo_eqs.py:65-78)
//   vl_cluster_kernel  update_mean_cluster (uqs update_eqs.py:55-71, lep update_eqs.py:90-112)
//                      + the sums update_alpha / elbo_data / elbo_cluster / update_c_and_d need
//   vl_scalar_kernel   alpha, mean_log_pi, a, b, (c, d), the ELBO and the loop state machine
//                      of run_cavi (cavi.py:120-169)
//
// The one-hot, quality-weighted read tensor E[n,l,b] (preparation.py:123-157) is never
// materialised: each thread builds its DMMA A-fragment in registers from the compact form
// (base code, log theta, log((1-theta)/(B-1))).  The other operand (responsibilities or
// haplotype table) is staged in shared memory by TMA bulk copies behind mbarriers.
#include "vl_device.cuh"
#include "vl_launch.h"
#include "../../include/viloca_b200.h"

namespace {

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// =========================================================================================
// update_mean_haplo: C[k,l,b] = sum_n (w_n z_nk) E[n,l,b], then softmax over b.
// CTA = 16 positions x all clusters; warp (lg, kh) = 8 positions x 5 bases x half of the
// cluster tiles; the DMMA M dimension is the position, so one thread ends up holding all
// five bases of its (position, cluster) pairs and the softmax over b is thread-local.
// =========================================================================================
template <int KT, int MODE>
__global__ void __launch_bounds__(VL_THREADS, 2)
vl_haplo_kernel(const VlProblem* __restrict__ probs, const int2* __restrict__ tiles) {
  constexpr int KTW = (KT + 1) / 2;
  constexpr int KS = vl_ks(KT);
  constexpr uint32_t STAGE_BYTES = VL_H_NC * KS * sizeof(double);
  extern __shared__ __align__(128) unsigned char vl_smem[];
  double* zs = reinterpret_cast<double*>(vl_smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(vl_smem + VL_STAGES * STAGE_BYTES);
  __shared__ double red[VL_THREADS / 32][4];

  const int2 tm = tiles[blockIdx.x];
  const VlProblem& P = probs[tm.x];
  const VlState* S = P.st;
  if (S->active == 0) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int lg = warp & 1, kh = warp >> 1, kt0 = kh * KTW;
  const int l = tm.y * VL_HT_POS + lg * 8 + g;
  const int Ls = P.Ls, Npad = P.Npad;
  const int nchunks = Npad / VL_H_NC;
  const double* __restrict__ zg = P.z;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < VL_STAGES; ++s) vl_mbar_init(&bars[s], 1);
    vl_fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < VL_STAGES - 1 && s < nchunks; ++s) {
      vl_mbar_expect_tx(&bars[s], STAGE_BYTES);
      vl_tma_load_1d(zs + s * VL_H_NC * KS, zg + (size_t)s * VL_H_NC * KS, STAGE_BYTES, &bars[s]);
    }
  }

  double acc[VL_B][KTW][2];
#pragma unroll
  for (int b = 0; b < VL_B; ++b)
#pragma unroll
    for (int i = 0; i < KTW; ++i) { acc[b][i][0] = 0.0; acc[b][i][1] = 0.0; }

  // compact operand of read n = 4*step + q at this thread's position, prefetched one step ahead
  const uint8_t* __restrict__ codes = P.codes;
  const double* __restrict__ ltp = P.lt;
  const double* __restrict__ lop = P.lo;
  const double* __restrict__ wp = P.w;
  int n = q;
  size_t off = (size_t)q * Ls + l;
  unsigned c_nx = codes[off];
  double w_nx = wp[n], lt_nx = 0.0, lo_nx = 0.0;
  if (MODE == VL_MODE_UQS) { lt_nx = ltp[off]; lo_nx = lop[off]; }

  for (int c = 0; c < nchunks; ++c) {
    const int st = c % VL_STAGES;
    if (tid == 0 && c + VL_STAGES - 1 < nchunks) {
      const int cn = c + VL_STAGES - 1, sn = cn % VL_STAGES;
      vl_mbar_expect_tx(&bars[sn], STAGE_BYTES);
      vl_tma_load_1d(zs + sn * VL_H_NC * KS, zg + (size_t)cn * VL_H_NC * KS, STAGE_BYTES, &bars[sn]);
    }
    vl_mbar_wait(&bars[st], (c / VL_STAGES) & 1);
    const double* zt = zs + st * VL_H_NC * KS + q * KS + g;
#pragma unroll
    for (int s = 0; s < VL_H_NC / 4; ++s) {
      const unsigned cc = c_nx;
      const double wv = w_nx, ltv = lt_nx, lov = lo_nx;
      n += 4;
      if (n < Npad) {
        off += (size_t)4 * Ls;
        c_nx = codes[off];
        w_nx = wp[n];
        if (MODE == VL_MODE_UQS) { lt_nx = ltp[off]; lo_nx = lop[off]; }
      }
      double a[VL_B];
      if (MODE == VL_MODE_UQS) {
        const double hit = (cc == VL_CODE_N) ? 0.0 : wv * ltv;
        const double base = (cc == VL_CODE_N) ? 0.0 : wv * lov;
#pragma unroll
        for (int b = 0; b < VL_B; ++b) a[b] = (cc == (unsigned)b) ? hit : base;
      } else {
#pragma unroll
        for (int b = 0; b < VL_B; ++b) a[b] = (cc == (unsigned)b) ? wv : 0.0;
      }
#pragma unroll
      for (int i = 0; i < KTW; ++i) {
        if (i < KTW - 1 || (KT % 2 == 0) || kh == 0) {
          const double bf = zt[(4 * s) * KS + (kt0 + i) * 8];
#pragma unroll
          for (int b = 0; b < VL_B; ++b) vl_dmma(acc[b][i], a[b], bf);
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue: + ref_part, softmax over b, the three sums, store mean_haplo ----
  const double mlg0 = S->mlg0, refoff = S->mlg1 - VL_LOG4;
  double b1 = 0.0, b2 = 0.0;
  if (MODE == VL_MODE_LEP) { b1 = S->mlt0; b2 = S->mlt1 - VL_LOG4; }
  const unsigned rc = P.ref[l];
  const bool lvalid = l < P.L;
  const int K = P.K;
  double s_ref = 0.0, s_nref = 0.0, s_ent = 0.0;
  double* hrow = P.h + (size_t)l * VL_B * KS;
#pragma unroll
  for (int i = 0; i < KTW; ++i) {
    if (i < KTW - 1 || (KT % 2 == 0) || kh == 0) {
      const int kb = (kt0 + i) * 8 + 2 * q;
      double out[VL_B][2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double lgt[VL_B];
        if (MODE == VL_MODE_UQS) {
#pragma unroll
          for (int b = 0; b < VL_B; ++b) lgt[b] = acc[b][i][j] + ((rc == (unsigned)b) ? mlg0 : refoff);
        } else {
          double tot = 0.0;
#pragma unroll
          for (int b = 0; b < VL_B; ++b) tot += acc[b][i][j];
#pragma unroll
          for (int b = 0; b < VL_B; ++b)
            lgt[b] = b1 * acc[b][i][j] + b2 * (tot - acc[b][i][j]) + ((rc == (unsigned)b) ? mlg0 : refoff);
        }
        double m = lgt[0];
#pragma unroll
        for (int b = 1; b < VL_B; ++b) m = fmax(m, lgt[b]);
        double e[VL_B], sum = 0.0;
#pragma unroll
        for (int b = 0; b < VL_B; ++b) { e[b] = exp(lgt[b] - m); sum += e[b]; }
        if (lvalid && (kb + j) < K) {
          const double ls = log(sum);
#pragma unroll
          for (int b = 0; b < VL_B; ++b) {
            const double hb = e[b] / sum;
            if (rc == (unsigned)b) s_ref += hb; else s_nref += hb;
            if (hb > 0.0) s_ent += hb * ((lgt[b] - m) - ls);
            out[b][j] = hb;
          }
        } else {
#pragma unroll
          for (int b = 0; b < VL_B; ++b) out[b][j] = 0.0;
        }
      }
#pragma unroll
      for (int b = 0; b < VL_B; ++b)
        *reinterpret_cast<double2*>(hrow + b * KS + kb) = make_double2(out[b][0], out[b][1]);
    }
  }
  s_ref = warp_sum(s_ref); s_nref = warp_sum(s_nref); s_ent = warp_sum(s_ent);
  if (lane == 0) { red[warp][0] = s_ref; red[warp][1] = s_nref; red[warp][2] = s_ent; }
  __syncthreads();
  if (tid < 3) {
    double v = 0.0;
#pragma unroll
    for (int wI = 0; wI < VL_THREADS / 32; ++wI) v += red[wI][tid];
    P.hpart[(size_t)tm.y * VL_HP_STRIDE + tid] = v;
  }
}

// =========================================================================================
// update_mean_cluster: X[n,k] = sum_{l,b} E[n,l,b] h[k,l,b], then softmax over k.
// CTA = 64 reads x all clusters; warp = 16 reads (two m8 tiles) x all cluster tiles.  The
// reduction index runs over (position, base): one k4 step covers 4 positions of one base.
// =========================================================================================
template <int KT, int MODE>
__global__ void __launch_bounds__(VL_THREADS, 2)
vl_cluster_kernel(const VlProblem* __restrict__ probs, const int2* __restrict__ tiles) {
  constexpr int KS = vl_ks(KT), K8 = vl_k8(KT);
  constexpr int STAGE_DOUBLES = VL_Z_PC * VL_B * KS;
  constexpr uint32_t STAGE_BYTES = STAGE_DOUBLES * sizeof(double);
  extern __shared__ __align__(128) unsigned char vl_smem[];
  double* hs = reinterpret_cast<double*>(vl_smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(vl_smem + VL_STAGES * STAGE_BYTES);
  double* red_cs = reinterpret_cast<double*>(vl_smem + VL_STAGES * STAGE_BYTES + 64);   // [4][K8]
  double* red_sc = red_cs + (VL_THREADS / 32) * K8;                                       // [4][4]

  const int2 tm = tiles[blockIdx.x];
  const VlProblem& P = probs[tm.x];
  const VlState* S = P.st;
  if (S->active == 0) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int Ls = P.Ls;
  const int nst = Ls / VL_Z_PC;
  const int n0 = tm.y * VL_ZT_READS + warp * 16 + g;
  const double* __restrict__ hg = P.h;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < VL_STAGES; ++s) vl_mbar_init(&bars[s], 1);
    vl_fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < VL_STAGES - 1 && s < nst; ++s) {
      vl_mbar_expect_tx(&bars[s], STAGE_BYTES);
      vl_tma_load_1d(hs + s * STAGE_DOUBLES, hg + (size_t)s * STAGE_DOUBLES, STAGE_BYTES, &bars[s]);
    }
  }

  double acc[2][KT][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) { acc[mt][kt][0] = 0.0; acc[mt][kt][1] = 0.0; }

  const uint8_t* __restrict__ codes = P.codes;
  const double* __restrict__ ltp = P.lt;
  const double* __restrict__ lop = P.lo;
  size_t off[2];
  off[0] = (size_t)n0 * Ls + q;
  off[1] = (size_t)(n0 + 8) * Ls + q;
  unsigned c_nx[2][2];
  double lt_nx[2][2], lo_nx[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int l4 = 0; l4 < 2; ++l4) {
      c_nx[mt][l4] = codes[off[mt] + 4 * l4];
      lt_nx[mt][l4] = 0.0; lo_nx[mt][l4] = 0.0;
      if (MODE == VL_MODE_UQS) { lt_nx[mt][l4] = ltp[off[mt] + 4 * l4]; lo_nx[mt][l4] = lop[off[mt] + 4 * l4]; }
    }

  for (int c = 0; c < nst; ++c) {
    const int st = c % VL_STAGES;
    if (tid == 0 && c + VL_STAGES - 1 < nst) {
      const int cn = c + VL_STAGES - 1, sn = cn % VL_STAGES;
      vl_mbar_expect_tx(&bars[sn], STAGE_BYTES);
      vl_tma_load_1d(hs + sn * STAGE_DOUBLES, hg + (size_t)cn * STAGE_DOUBLES, STAGE_BYTES, &bars[sn]);
    }
    unsigned cc[2][2];
    double hitv[2][2], offv[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int l4 = 0; l4 < 2; ++l4) {
        cc[mt][l4] = c_nx[mt][l4];
        if (MODE == VL_MODE_UQS) {
          hitv[mt][l4] = (cc[mt][l4] == VL_CODE_N) ? 0.0 : lt_nx[mt][l4];
          offv[mt][l4] = (cc[mt][l4] == VL_CODE_N) ? 0.0 : lo_nx[mt][l4];
        } else {
          hitv[mt][l4] = 1.0; offv[mt][l4] = 0.0;
        }
      }
    if (c + 1 < nst) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        off[mt] += VL_Z_PC;
#pragma unroll
        for (int l4 = 0; l4 < 2; ++l4) {
          c_nx[mt][l4] = codes[off[mt] + 4 * l4];
          if (MODE == VL_MODE_UQS) { lt_nx[mt][l4] = ltp[off[mt] + 4 * l4]; lo_nx[mt][l4] = lop[off[mt] + 4 * l4]; }
        }
      }
    }
    vl_mbar_wait(&bars[st], (c / VL_STAGES) & 1);
    const double* ht = hs + st * STAGE_DOUBLES + q * VL_B * KS + g;
#pragma unroll
    for (int l4 = 0; l4 < 2; ++l4) {
#pragma unroll
      for (int b = 0; b < VL_B; ++b) {
        const double a0 = (cc[0][l4] == (unsigned)b) ? hitv[0][l4] : offv[0][l4];
        const double a1 = (cc[1][l4] == (unsigned)b) ? hitv[1][l4] : offv[1][l4];
        const double* hb = ht + (4 * l4 * VL_B + b) * KS;
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
          const double bf = hb[kt * 8];
          vl_dmma(acc[0][kt], a0, bf);
          vl_dmma(acc[1][kt], a1, bf);
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue: + mean_log_pi, softmax over k, the reductions, store mean_cluster ----
  const double* __restrict__ mlp = P.mlp;
  const int K = P.K, N = P.N;
  double b1 = 0.0, b2 = 0.0;
  const double Lf = (double)P.L;
  if (MODE == VL_MODE_LEP) { b1 = S->mlt0; b2 = S->mlt1 - VL_LOG4; }
  double s_data = 0.0, s_ent = 0.0, s_c = 0.0, s_d = 0.0;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int n = n0 + 8 * mt;
    const bool valid = n < N;
    const double wn = P.w[n];
    const double mc = P.mcount[n];
    double t[KT][2];
    double m = neg_inf();
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = kt * 8 + 2 * q + j;
        const double x = acc[mt][kt][j];
        double tv;
        if (k < K) {
          const double lp = mlp[k];
          tv = (MODE == VL_MODE_UQS) ? (x + lp) : (b1 * x + b2 * (Lf - x) + lp);
        } else {
          tv = neg_inf();
        }
        t[kt][j] = tv;
        m = fmax(m, tv);
      }
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
    double sum = 0.0, u = 0.0, v = 0.0;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = kt * 8 + 2 * q + j;
        double e = 0.0;
        if (k < K) {
          const double dt = t[kt][j] - m;
          e = exp(dt);
          if (e > 0.0) u += e * dt;
          v += e * acc[mt][kt][j];
        }
        t[kt][j] = e;
        sum += e;
      }
    const double psum = sum;   // this thread's share of the row sum
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    if (valid) {
      // sum_k z log z = (sum_k e_k (t_k - m)) / s - log s   (0 log 0 := 0)
      s_ent += u / sum - ((q == 0) ? log(sum) : 0.0);
      if (MODE == VL_MODE_UQS) {
        s_data += wn * (v / sum);
      } else {
        s_c += wn * (v / sum);
        s_d += wn * ((mc * psum - v) / sum);
      }
    }
    double* zrow = P.z + (size_t)n * KS + 2 * q;
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      const double z0 = valid ? t[kt][0] / sum : 0.0;
      const double z1 = valid ? t[kt][1] / sum : 0.0;
      *reinterpret_cast<double2*>(zrow + kt * 8) = make_double2(z0, z1);
      // the weighted column sums of this thread's two rows are collected in acc[0]
      if (mt == 0) { acc[0][kt][0] = wn * z0; acc[0][kt][1] = wn * z1; }
      else { acc[0][kt][0] += wn * z0; acc[0][kt][1] += wn * z1; }
    }
  }
#pragma unroll
  for (int kt = 0; kt < KT; ++kt)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      double v = acc[0][kt][j];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (g == 0) red_cs[warp * K8 + kt * 8 + 2 * q + j] = v;
    }
  s_data = warp_sum(s_data); s_ent = warp_sum(s_ent); s_c = warp_sum(s_c); s_d = warp_sum(s_d);
  if (lane == 0) {
    red_sc[warp * 4 + 0] = s_data; red_sc[warp * 4 + 1] = s_ent;
    red_sc[warp * 4 + 2] = s_c;    red_sc[warp * 4 + 3] = s_d;
  }
  __syncthreads();
  double* zp = P.zpart + (size_t)tm.y * (K8 + VL_ZP_EXTRA);
  for (int k = tid; k < K8 + 4; k += VL_THREADS) {
    double v = 0.0;
    if (k < K8) {
#pragma unroll
      for (int wI = 0; wI < VL_THREADS / 32; ++wI) v += red_cs[wI * K8 + k];
    } else {
#pragma unroll
      for (int wI = 0; wI < VL_THREADS / 32; ++wI) v += red_sc[wI * 4 + (k - K8)];
    }
    zp[k] = v;
  }
}

// =========================================================================================
// Scalar update, ELBO, and the loop state machine of run_cavi.
// =========================================================================================
__global__ void __launch_bounds__(256)
vl_scalar_kernel(const VlProblem* __restrict__ probs, int* __restrict__ done_count) {
  const VlProblem& P = probs[blockIdx.x];
  VlState* S = P.st;
  if (S->active == 0) return;
  __shared__ double sh_cs[VL_MAX_CLUSTERS];
  __shared__ double sh_sc[8];
  __shared__ double sh_t[6][VL_MAX_CLUSTERS];
  __shared__ double sh_sum[6];
  const int tid = threadIdx.x;
  const int K = P.K, K8 = 8 * P.KT, ZS = K8 + VL_ZP_EXTRA;

  if (tid < K8 + 4) {
    // fixed-order sum over the read tiles: column sums (k < K8) and the four scalars
    const double* zp = P.zpart + tid;   // column tid, or scalar tid - K8 stored right after the columns
    double v = 0.0;
    for (int t = 0; t < P.n_ztiles; ++t) v += zp[(size_t)t * ZS];
    if (tid < K8) sh_cs[tid] = v; else sh_sc[tid - K8] = v;
  } else if (tid < K8 + 7) {
    const int j = tid - K8 - 4;
    double v = 0.0;
    for (int t = 0; t < P.n_htiles; ++t) v += P.hpart[(size_t)t * VL_HP_STRIDE + j];
    sh_sc[4 + j] = v;
  }
  __syncthreads();

  const double alpha0 = S->alpha0;
  if (tid < K) {
    const double cs = sh_cs[tid];
    const double al = alpha0 + cs;                       // update_alpha, update_eqs.py:114-121
    const double lp = vl_digamma(al) - S->dsum_alpha;    // get_mean_log_pi, update_eqs.py:40-46
    P.alpha[tid] = al;
    P.mlp[tid] = lp;
    sh_t[0][tid] = (alpha0 - 1.0) * lp;
    sh_t[1][tid] = (al - 1.0) * lp;
    sh_t[2][tid] = lgamma(al);
    sh_t[3][tid] = al;
    sh_t[4][tid] = cs * lp;
  }
  __syncthreads();
  if (tid < 5) {
    double v = 0.0;
    for (int k = 0; k < K; ++k) v += sh_t[tid][k];
    sh_sum[tid] = v;
  }
  __syncthreads();
  if (tid != 0) return;

  const int mode = P.mode;
  const double z_data = sh_sc[0], z_ent = sh_sc[1], z_c = sh_sc[2], z_d = sh_sc[3];
  const double h_ref = sh_sc[4], h_nref = sh_sc[5], h_ent = sh_sc[6];
  const double sum_a0m1_lp = sh_sum[0], sum_am1_lp = sh_sum[1], sum_lgamma = sh_sum[2];
  const double sum_alpha = sh_sum[3], sum_cs_lp = sh_sum[4];

  // update_a_and_b + get_mean_log_beta_dist (update_eqs.py:103-112, 48-53)
  const double a = S->a0 + h_ref, b = S->b0 + h_nref;
  const double mlg0 = vl_digamma(a) - S->dsum_ab, mlg1 = vl_digamma(b) - S->dsum_ab;
  double c = S->c0, d = S->d0, mlt0 = S->mlt0, mlt1 = S->mlt1;
  if (mode == VL_MODE_LEP) {                             // update_c_and_d, lep update_eqs.py:174-188
    c = S->c0 + z_c; d = S->d0 + z_d;
    mlt0 = vl_digamma(c) - S->dsum_cd; mlt1 = vl_digamma(d) - S->dsum_cd;
  }
  // ELBO (elbo_eqs.py:27-31; lep elbo_eqs.py:34-39), terms added in the reference's order
  const double t_data = (mode == VL_MODE_UQS) ? z_data : (mlt0 * z_c + (mlt1 - VL_LOG4) * z_d);
  const double lnB_alpha = sum_lgamma - lgamma(sum_alpha);
  const double t_pi = ((-1.0) * S->lnB0 + sum_a0m1_lp) - ((-1.0) * lnB_alpha + sum_am1_lp);
  const double t_cluster = sum_cs_lp - z_ent;
  const double t_haplo = (mlg0 * h_ref + (mlg1 - VL_LOG4) * h_nref) - h_ent;
  const double t_gamma = ((S->a0 - 1.0) * mlg0 + (S->b0 - 1.0) * mlg1 - S->betaln_ab0) -
                         ((a - 1.0) * mlg0 + (b - 1.0) * mlg1 - vl_betaln(a, b));
  double t_theta = 0.0;
  double elbo = t_data;
  elbo += t_pi; elbo += t_cluster; elbo += t_haplo; elbo += t_gamma;
  if (mode == VL_MODE_LEP) {
    t_theta = ((S->c0 - 1.0) * mlt0 + (S->d0 - 1.0) * mlt1 - S->betaln_cd0) -
              ((c - 1.0) * mlt0 + (d - 1.0) * mlt1 - vl_betaln(c, d));
    elbo += t_theta;
  }
  S->a = a; S->b = b; S->c = c; S->d = d;
  S->mlg0 = mlg0; S->mlg1 = mlg1; S->mlt0 = mlt0; S->mlt1 = mlt1;
  S->elbo = elbo; S->t_data = t_data; S->t_pi = t_pi; S->t_cluster = t_cluster;
  S->t_haplo = t_haplo; S->t_gamma = t_gamma; S->t_theta = t_theta;

  // ---- loop state machine: uqs cavi.py:146-169, lep cavi.py:167-199 ----
  int iter = S->iter, converged = S->converged, status = VL_STATUS_RUNNING;
  int n_hist = S->n_hist;
  double h_last = S->h_last, h_prev = S->h_prev;
  const bool append = (mode == VL_MODE_UQS) || ((iter & 1) == 0);
  if (append) {
    if (n_hist < S->max_iter) P.hist[n_hist] = elbo;
    h_prev = h_last; h_last = elbo; ++n_hist;
  }
  const int n_updates = S->n_updates + 1;
  bool stop = false;
  if (iter > 1) {
    const double dec_tol = (mode == VL_MODE_UQS) ? 1e-5 : 1e-8;
    const double thr = (mode == VL_MODE_UQS) ? S->thr : 1e-3;
    if (isnan(elbo)) { status = VL_STATUS_NAN; stop = true; }
    else if (h_prev > elbo && fabs(elbo - h_prev) > dec_tol) { status = VL_STATUS_DECREASING; stop = true; }
    else if (fabs(elbo - h_prev) < thr) { converged = 1; if (mode == VL_MODE_LEP) iter += 1; }
    else if (mode == VL_MODE_LEP) { iter = 0; }
  }
  if (!stop) {
    iter += 1;
    if (converged && iter >= 10) { status = VL_STATUS_CONVERGED; stop = true; }
    else if (n_updates >= S->max_iter) { status = VL_STATUS_MAXITER; stop = true; }
  }
  if (!stop && iter <= 1) {   // refresh of the digamma sums for the next update, cavi.py:122-128
    S->dsum_alpha = vl_digamma(sum_alpha);
    S->dsum_ab = vl_digamma(a + b);
    if (mode == VL_MODE_LEP) S->dsum_cd = vl_digamma(c + d);
  }
  S->iter = iter; S->converged = converged; S->n_hist = n_hist; S->n_updates = n_updates;
  S->h_last = h_last; S->h_prev = h_prev; S->status = status;
  if (stop) {
    S->active = 0;
    atomicAdd(done_count, 1);
  }
}

// One thread per problem: state_init_dict of draw_init_state (initialization.py:6-43) and the
// first refresh of the digamma sums (cavi.py:122-128 at iter = 0).
__global__ void vl_init_kernel(const VlProblem* __restrict__ probs, int n_probs) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_probs) return;
  const VlProblem& P = probs[p];
  VlState* S = P.st;
  const int K = P.K, K8 = 8 * P.KT;
  const double alpha0 = S->alpha0;
  double sum_alpha = 0.0;
  for (int k = 0; k < K; ++k) sum_alpha += alpha0;
  const double ds = vl_digamma(sum_alpha);
  const double lp0 = vl_digamma(alpha0) - ds;
  for (int k = 0; k < K8; ++k) { P.alpha[k] = (k < K) ? alpha0 : 0.0; P.mlp[k] = (k < K) ? lp0 : 0.0; }
  S->lnB0 = K * lgamma(alpha0) - lgamma(sum_alpha);
  S->betaln_ab0 = vl_betaln(S->a0, S->b0);
  S->a = S->a0; S->b = S->b0; S->c = S->c0; S->d = S->d0;
  S->dsum_alpha = ds;
  S->dsum_ab = vl_digamma(S->a0 + S->b0);
  S->betaln_cd0 = 0.0; S->dsum_cd = 0.0;
  if (P.mode == VL_MODE_LEP) {
    S->betaln_cd0 = vl_betaln(S->c0, S->d0);
    S->dsum_cd = vl_digamma(S->c0 + S->d0);
  }
  S->elbo = 0.0; S->t_data = S->t_pi = S->t_cluster = S->t_haplo = S->t_gamma = S->t_theta = 0.0;
  S->h_last = 0.0; S->h_prev = 0.0;
  S->iter = 0; S->n_updates = 0; S->n_hist = 0; S->converged = 0;
  S->status = VL_STATUS_RUNNING; S->active = 1;
}

// number of called (non-N) positions per read: n_non_N of lep preparation.py:13
__global__ void vl_mcount_kernel(const uint8_t* __restrict__ codes, double* __restrict__ mcount,
                                 int Npad, int Ls) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= Npad) return;
  const int lane = threadIdx.x & 31;
  int cnt = 0;
  for (int l = lane; l < Ls; l += 32) cnt += (codes[(size_t)n * Ls + l] != VL_CODE_N);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) mcount[n] = (double)cnt;
}

// mean_haplo [Ls][B][KS] -> reference layout (K, L, B)
__global__ void vl_export_haplo_kernel(const double* __restrict__ h, double* __restrict__ out,
                                       int K, int L, int KS) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)K * L * VL_B;
  if (i >= total) return;
  const int b = (int)(i % VL_B);
  const int l = (int)((i / VL_B) % L);
  const int k = (int)(i / ((size_t)VL_B * L));
  out[i] = h[((size_t)l * VL_B + b) * KS + k];
}

// merge_cluster_assignments (analyze_results.py:160-167): members added in ascending k
__global__ void vl_merge_kernel(const double* __restrict__ z, double* __restrict__ zm,
                                const int* __restrict__ group_of, int N, int K, int G, int KS) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N * KS) return;
  const int n = (int)(i / KS), j = (int)(i % KS);
  double s = 0.0;
  if (j < G)
    for (int k = 0; k < K; ++k) if (group_of[k] == j) s += z[(size_t)n * KS + k];
  zm[i] = s;
}

template <int KT, int MODE>
cudaError_t launch_pair(const VlProblem* probs, const int2* htiles, int n_ht, const int2* ztiles,
                        int n_zt, bool do_h, bool do_z, cudaStream_t stream) {
  constexpr int KS = vl_ks(KT), K8 = vl_k8(KT);
  const size_t smem_h = (size_t)VL_STAGES * VL_H_NC * KS * 8 + 64;
  const size_t smem_z = (size_t)VL_STAGES * VL_Z_PC * VL_B * KS * 8 + 64 + (size_t)(VL_THREADS / 32) * (K8 + 4) * 8;
  // per-device attribute; cheap enough to set on every launch (contexts on several GPUs may
  // live in one process)
  cudaError_t e = cudaFuncSetAttribute(vl_haplo_kernel<KT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_h);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(vl_cluster_kernel<KT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_z);
  if (e != cudaSuccess) return e;
  if (do_h && n_ht > 0) vl_haplo_kernel<KT, MODE><<<n_ht, VL_THREADS, smem_h, stream>>>(probs, htiles);
  if (do_z && n_zt > 0) vl_cluster_kernel<KT, MODE><<<n_zt, VL_THREADS, smem_z, stream>>>(probs, ztiles);
  return cudaGetLastError();
}

}  // namespace

int vl_supported_kt(int K) {
  const int kts[] = {2, 4, 7, 10, 13, 16};
  for (int kt : kts) if (K <= 8 * kt) return kt;
  return -1;
}

cudaError_t vl_launch_contractions(int kt, int mode, const VlProblem* probs, const int2* htiles,
                                   int n_ht, const int2* ztiles, int n_zt, bool do_h, bool do_z,
                                   cudaStream_t stream) {
#define VL_CASE(KT_)                                                                               \
  case KT_:                                                                                        \
    return (mode == VL_MODE_UQS)                                                                   \
               ? launch_pair<KT_, VL_MODE_UQS>(probs, htiles, n_ht, ztiles, n_zt, do_h, do_z, stream) \
               : launch_pair<KT_, VL_MODE_LEP>(probs, htiles, n_ht, ztiles, n_zt, do_h, do_z, stream);
  switch (kt) {
    VL_CASE(2) VL_CASE(4) VL_CASE(7) VL_CASE(10) VL_CASE(13) VL_CASE(16)
    default: return cudaErrorInvalidValue;
  }
#undef VL_CASE
}

cudaError_t vl_launch_scalar(const VlProblem* probs, int n_probs, int* done_count, cudaStream_t stream) {
  if (n_probs > 0) vl_scalar_kernel<<<n_probs, 256, 0, stream>>>(probs, done_count);
  return cudaGetLastError();
}

cudaError_t vl_launch_init(const VlProblem* probs, int n_probs, cudaStream_t stream) {
  if (n_probs > 0) vl_init_kernel<<<(n_probs + 127) / 128, 128, 0, stream>>>(probs, n_probs);
  return cudaGetLastError();
}

cudaError_t vl_launch_mcount(const uint8_t* codes, double* mcount, int Npad, int Ls, cudaStream_t stream) {
  vl_mcount_kernel<<<(Npad + 3) / 4, 128, 0, stream>>>(codes, mcount, Npad, Ls);
  return cudaGetLastError();
}

cudaError_t vl_launch_export_haplo(const double* h, double* out, int K, int L, int KS, cudaStream_t stream) {
  const size_t total = (size_t)K * L * VL_B;
  vl_export_haplo_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(h, out, K, L, KS);
  return cudaGetLastError();
}

cudaError_t vl_launch_merge(const double* z, double* zm, const int* group_of, int N, int K, int G,
                            int KS, cudaStream_t stream) {
  const size_t total = (size_t)N * KS;
  vl_merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(z, zm, group_of, N, K, G, KS);
  return cudaGetLastError();
}
