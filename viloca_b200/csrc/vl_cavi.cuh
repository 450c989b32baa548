// Device code shared by the general-path kernels (vl_kernels.cu) and the resident kernel
// (vl_resident.cu): small helpers and the scalar part of one CAVI update.
#pragma once
#include <type_traits>

#include "vl_device.cuh"
#include "../../include/viloca_b200.h"

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exp(x) for x <= 0 (the only arguments a max-subtracted softmax produces; -inf allowed), N
// values at once.  Branch-free, so that the N evaluations are N independent instruction
// streams: FP64 latency, not throughput, bounds the softmax epilogues when few warps run
// (CUDA's exp() is a serial chain of ~25 dependent FP64 operations per value).
// Cody-Waite reduction x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial (truncation
// 4e-18), scaling by 2^k in two steps so that results in the subnormal range are rounded once,
// like a correctly rounded exp: exp(x) == 0 exactly for x < -745.14 (the zero pattern of
// mean_cluster is part of the parity contract).  Max error about 1 ulp.
template <int N>
__device__ __forceinline__ void vl_exp_nonpos(const double (&x)[N], double (&out)[N]) {
  constexpr double L2E = 1.4426950408889634074, MAGIC = 6755399441055744.0;   // 1.5 * 2^52
  constexpr double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  double r[N], p[N];
  int k[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double xi = fmax(x[i], -800.0);
    const double t = fma(xi, L2E, MAGIC);
    k[i] = __double2loint(t);
    const double kk = t - MAGIC;
    r[i] = fma(kk, -LN2_LO, fma(kk, -LN2_HI, xi));
    p[i] = 1.0 / 6227020800.0;
  }
  constexpr double C[13] = {1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0,
                            1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0};
#pragma unroll
  for (int j = 0; j < 13; ++j)
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = fma(p[i], r[i], C[j]);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const int k1 = k[i] >> 1, k2 = k[i] - k1;
    const double s1 = __hiloint2double((k1 + 1023) << 20, 0), s2 = __hiloint2double((k2 + 1023) << 20, 0);
    out[i] = (p[i] * s1) * s2;
  }
}

// Run f(std::integral_constant<int, kt>) for the run-time tile count kt in [1, KTM].  The DMMA
// loops are instantiated per tile count because a predicated-off DMMA still occupies the FP64
// pipe for a full issue slot on sm_100a (measured: 2.2x on the kt = 1 loops).
template <int KTM, int KT = 1, class F>
__device__ __forceinline__ void vl_dispatch_tiles(int kt, F&& f) {
  if constexpr (KT >= KTM) {
    f(std::integral_constant<int, KTM>{});
  } else {
    if (kt == KT) f(std::integral_constant<int, KT>{});
    else vl_dispatch_tiles<KTM, KT + 1>(kt, f);
  }
}

// Cooperative copy of a restart's state into shared memory with coherent (L2) loads: one memory
// latency instead of one per field.  The caller synchronises before reading dst.
__device__ __forceinline__ void vl_load_state(VlState* dst, const VlState* src, int tid) {
  static_assert(sizeof(VlState) % 8 == 0, "state is copied in 8-byte words");
  constexpr int W = sizeof(VlState) / 8;
  if (tid < W)
    reinterpret_cast<unsigned long long*>(dst)[tid] = __ldcg(reinterpret_cast<const unsigned long long*>(src) + tid);
}

// sum_{t<n} p[t * stride] in index order, sixteen loads in flight (adding the 0.0 of a padded slot
// changes nothing, so the value equals the plain sequential loop)
__device__ __forceinline__ double vl_ordered_sum(const double* p, size_t stride, int n) {
  double v = 0.0;
#ifndef VL_SUM_WIDTH
#define VL_SUM_WIDTH 16
#endif
  for (int t = 0; t < n; t += VL_SUM_WIDTH) {
    double x[VL_SUM_WIDTH];
#pragma unroll
    for (int u = 0; u < VL_SUM_WIDTH; ++u) x[u] = (t + u < n) ? __ldcg(p + (size_t)(t + u) * stride) : 0.0;
#pragma unroll
    for (int u = 0; u < VL_SUM_WIDTH; ++u) v += x[u];
  }
  return v;
}

// Named barrier over the first `n` threads of the CTA (n a multiple of 32).  Barrier 0 over
// the whole CTA is __syncthreads().
template <int BAR>
__device__ __forceinline__ void vl_bar_sync(int n) {
  asm volatile("bar.sync %0, %1;\n" :: "r"(BAR), "r"(n) : "memory");
}

// Shared-memory scratch of vl_scalar_update.  The caller fills cs (column sums of w*z per
// slot of the current layout) and sc (sum w z X | sum z log z | lep c | lep d | sum h ref |
// sum h (1-ref) | sum h log h) before the call.
struct VlScalarSmem {
  double cs[VL_MAX_CLUSTERS];
  double sc[8];
  double part[VL_MAX_CLUSTERS / 32][5];
  int slot[VL_MAX_CLUSTERS];
  int nlive[VL_MAX_CLUSTERS / 32], fdead[VL_MAX_CLUSTERS / 32];
  double fn[10];     // digamma / lgamma of the updated Beta parameters, one thread each
  int stop;
};

// alpha, mean_log_pi, a, b, (c, d), the ELBO in the reference's term order, the loop state
// machine of run_cavi (uqs cavi.py:120-169, lep cavi.py:135-199 as written) and the layout of
// the next iteration.  Called by threads 0..127 of the CTA (one per cluster), which
// synchronise on named barrier BAR; returns 1 to all of them when the restart left its loop.
// R is the state as it was before the update (a shared-memory copy where latency matters), S the
// state in global memory that receives the results.
// development aid (-DVL_RS_PROFILE in vl_kernels.cu): cycles of restart 0's scalar update by phase
#ifdef VL_SU_PROFILE
#define SU_TICK(slot)                                                                          \
  do {                                                                                         \
    if (su_on_) {                                                                              \
      const long long n_ = clock64();                                                          \
      atomicAdd(&vl_su_cycles[slot], (unsigned long long)(n_ - su_t_));                        \
      su_t_ = clock64();                                                                       \
    }                                                                                          \
  } while (0)
#else
#define SU_TICK(slot) do { } while (0)
#endif

template <int BAR>
__device__ __forceinline__ int vl_scalar_update(const VlProblem& P, const VlState& R, VlState* S, VlScalarSmem& sm,
                                                int tid, int* done_count, int4* pstat_entry) {
  constexpr int NT = VL_MAX_CLUSTERS;
  const int warp = tid >> 5, lane = tid & 31;
  const int K = P.K;
  const int ktl = R.ktl, nlay = R.nlay, ghost = R.ghost;
#ifdef VL_SU_PROFILE
  const bool su_on_ = P.window == 0 && P.restart == 0 && tid == 0;
  long long su_t_ = clock64();
#endif

  // z and h are now stored in the current layout: remember it for the exports
  const int old_lay = (tid < 8 * ktl) ? P.lay[tid] : -1;
  if (tid < 8 * ktl) P.lay_z[tid] = old_lay;
  if (tid < K) sm.slot[tid] = ghost;
  vl_bar_sync<BAR>(NT);
  if (tid < nlay && tid != ghost) sm.slot[old_lay] = tid;
  vl_bar_sync<BAR>(NT);
  SU_TICK(0);

  // the special functions of the updated Beta parameters depend only on sm.sc: the last threads
  // of the CTA evaluate one each while the others work on their clusters
  {
    const int j = NT - 1 - tid;
    if (j < ((P.mode == VL_MODE_LEP) ? 10 : 5)) {
      const double a = R.a0 + sm.sc[4], b = R.b0 + sm.sc[5];
      const double c = R.c0 + sm.sc[2], d = R.d0 + sm.sc[3];
      // one code path for all of them (a switch over ten different calls would serialise the warp)
      const double arg = (j == 0 || j == 2) ? a : (j == 1 || j == 3) ? b : (j == 4) ? a + b
                       : (j == 5 || j == 7) ? c : (j == 6 || j == 8) ? d : c + d;
      double psi, lg;
      vl_psi_lgamma(arg, psi, lg);
      const double v = (j == 0 || j == 1 || j == 5 || j == 6) ? psi : lg;
      sm.fn[j] = v;
    }
  }

  const double alpha0 = R.alpha0;
  double cs = 0.0, lp = 0.0, q0 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0, q4 = 0.0;
  int my_slot = 0;
  if (tid < K) {
    my_slot = sm.slot[tid];
    cs = sm.cs[my_slot];                                 // a dead cluster has the ghost's sum
    const double al = alpha0 + cs;                       // update_alpha, update_eqs.py:114-121
    double psi_al, lg_al;
    vl_psi_lgamma(al, psi_al, lg_al);
    lp = psi_al - R.dsum_alpha;                          // get_mean_log_pi, update_eqs.py:40-46
    P.alpha[tid] = al;
    P.mlp[tid] = lp;
    q0 = (alpha0 - 1.0) * lp; q1 = (al - 1.0) * lp; q2 = lg_al; q3 = al; q4 = cs * lp;
  }
  q0 = warp_sum(q0); q1 = warp_sum(q1); q2 = warp_sum(q2); q3 = warp_sum(q3); q4 = warp_sum(q4);
  if (lane == 0) { sm.part[warp][0] = q0; sm.part[warp][1] = q1; sm.part[warp][2] = q2; sm.part[warp][3] = q3; sm.part[warp][4] = q4; }
  // candidate layout of the next iteration (used unless the loop stops): a cluster stays while
  // any read gives it a non-zero responsibility (column sum > 0; the weights are positive).
  // All others have mean_cluster[:,k] == 0 exactly, hence identical mean_haplo rows, logits and
  // mean_log_pi, and are represented by the single ghost slot with multiplicity.
  const bool is_cluster = tid < K;
  const bool live = is_cluster && (R.skip == 0 || cs > 0.0);
  const unsigned live_mask = __ballot_sync(0xffffffffu, live);
  const unsigned dead_mask = __ballot_sync(0xffffffffu, is_cluster && !live);
  if (lane == 0) {
    sm.nlive[warp] = __popc(live_mask);
    sm.fdead[warp] = dead_mask ? (warp * 32 + __ffs(dead_mask) - 1) : -1;
  }
  SU_TICK(1);
  vl_bar_sync<BAR>(NT);
  SU_TICK(2);

  if (tid == 0) {
    double sum5[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      double v = 0.0;
#pragma unroll
      for (int wI = 0; wI < NT / 32; ++wI) v += sm.part[wI][j];
      sum5[j] = v;
    }
    const int mode = P.mode;
    const double z_data = sm.sc[0], z_ent = sm.sc[1], z_c = sm.sc[2], z_d = sm.sc[3];
    const double h_ref = sm.sc[4], h_nref = sm.sc[5], h_ent = sm.sc[6];
    const double sum_a0m1_lp = sum5[0], sum_am1_lp = sum5[1], sum_lgamma = sum5[2];
    const double sum_alpha = sum5[3], sum_cs_lp = sum5[4];

    // update_a_and_b + get_mean_log_beta_dist (update_eqs.py:103-112, 48-53)
    const double a = R.a0 + h_ref, b = R.b0 + h_nref;
    const double mlg0 = sm.fn[0] - R.dsum_ab, mlg1 = sm.fn[1] - R.dsum_ab;
    double c = R.c0, d = R.d0, mlt0 = R.mlt0, mlt1 = R.mlt1;
    if (mode == VL_MODE_LEP) {                             // update_c_and_d, lep update_eqs.py:174-188
      c = R.c0 + z_c; d = R.d0 + z_d;
      mlt0 = sm.fn[5] - R.dsum_cd; mlt1 = sm.fn[6] - R.dsum_cd;
    }
    // ELBO (elbo_eqs.py:27-31; lep elbo_eqs.py:34-39), terms added in the reference's order
    const double t_data = (mode == VL_MODE_UQS) ? z_data : (mlt0 * z_c + (mlt1 - VL_LOG4) * z_d);
    const double lnB_alpha = sum_lgamma - vl_lgamma(sum_alpha);
    const double t_pi = ((-1.0) * R.lnB0 + sum_a0m1_lp) - ((-1.0) * lnB_alpha + sum_am1_lp);
    const double t_cluster = sum_cs_lp - z_ent;
    const double t_haplo = (mlg0 * h_ref + (mlg1 - VL_LOG4) * h_nref) - h_ent;
    const double t_gamma = ((R.a0 - 1.0) * mlg0 + (R.b0 - 1.0) * mlg1 - R.betaln_ab0) -
                           ((a - 1.0) * mlg0 + (b - 1.0) * mlg1 - (sm.fn[2] + sm.fn[3] - sm.fn[4]));
    double t_theta = 0.0;
    double elbo = t_data;
    elbo += t_pi; elbo += t_cluster; elbo += t_haplo; elbo += t_gamma;
    if (mode == VL_MODE_LEP) {
      t_theta = ((R.c0 - 1.0) * mlt0 + (R.d0 - 1.0) * mlt1 - R.betaln_cd0) -
                ((c - 1.0) * mlt0 + (d - 1.0) * mlt1 - (sm.fn[7] + sm.fn[8] - sm.fn[9]));
      elbo += t_theta;
    }
    S->a = a; S->b = b; S->c = c; S->d = d;
    S->mlg0 = mlg0; S->mlg1 = mlg1; S->mlt0 = mlt0; S->mlt1 = mlt1;
    S->elbo = elbo; S->t_data = t_data; S->t_pi = t_pi; S->t_cluster = t_cluster;
    S->t_haplo = t_haplo; S->t_gamma = t_gamma; S->t_theta = t_theta;

    // ---- loop state machine: uqs cavi.py:146-169, lep cavi.py:167-199 ----
    int iter = R.iter, converged = R.converged, status = VL_STATUS_RUNNING;
    int n_hist = R.n_hist;
    double h_last = R.h_last, h_prev = R.h_prev;
    const bool append = (mode == VL_MODE_UQS) || ((iter & 1) == 0);
    if (append) {
      if (n_hist < R.max_iter) P.hist[n_hist] = elbo;
      h_prev = h_last; h_last = elbo; ++n_hist;
    }
    const int n_updates = R.n_updates + 1;
    bool stop = false;
    if (iter > 1) {
      const double dec_tol = (mode == VL_MODE_UQS) ? 1e-5 : 1e-8;
      const double thr = (mode == VL_MODE_UQS) ? R.thr : 1e-3;
      if (isnan(elbo)) { status = VL_STATUS_NAN; stop = true; }
      else if (h_prev > elbo && fabs(elbo - h_prev) > dec_tol) { status = VL_STATUS_DECREASING; stop = true; }
      else if (fabs(elbo - h_prev) < thr) { converged = 1; if (mode == VL_MODE_LEP) iter += 1; }
      else if (mode == VL_MODE_LEP) { iter = 0; }
    }
    if (!stop) {
      iter += 1;
      if (converged && iter >= 10) { status = VL_STATUS_CONVERGED; stop = true; }
      else if (n_updates >= R.max_iter) { status = VL_STATUS_MAXITER; stop = true; }
    }
    if (!stop && iter <= 1) {   // refresh of the digamma sums for the next update, cavi.py:122-128
      S->dsum_alpha = vl_digamma(sum_alpha);
      S->dsum_ab = vl_digamma(a + b);
      if (mode == VL_MODE_LEP) S->dsum_cd = vl_digamma(c + d);
    }
    S->iter = iter; S->converged = converged; S->n_hist = n_hist; S->n_updates = n_updates;
    S->h_last = h_last; S->h_prev = h_prev; S->status = status;
    S->nlay_z = nlay; S->ghost_z = ghost; S->ktl_z = ktl;

    int nl = nlay, kt_new = ktl;
    if (!stop) {
      int n_live = 0;
#pragma unroll
      for (int wI = 0; wI < NT / 32; ++wI) n_live += sm.nlive[wI];
      const int n_dead = K - n_live;
      nl = n_live + (n_dead > 0 ? 1 : 0);
      kt_new = (nl + 7) / 8;
      S->ghost_mult = (double)n_dead;
      S->nlay = nl; S->ghost = (n_dead > 0) ? n_live : -1;
      S->ktl = kt_new;
      S->ktring[n_updates & 1] = kt_new;      // update number n_updates (the next one) stores z in this layout
    } else {
      S->active = 0;
      atomicAdd(done_count, 1);
    }
    *pstat_entry = make_int4(stop ? 0 : 1, max(kt_new, ktl), nl, n_updates);
    sm.stop = stop ? 1 : 0;
  }
  SU_TICK(3);
  vl_bar_sync<BAR>(NT);
  if (sm.stop) {
    if (tid == 0) { __threadfence(); *reinterpret_cast<volatile int*>(&S->pub) = R.n_updates + 1; }
    return 1;
  }

  // ---- write the layout of the next iteration ----
  int n_live = 0, first_dead = -1, before = 0;
#pragma unroll
  for (int wI = 0; wI < NT / 32; ++wI) {
    if (wI < warp) before += sm.nlive[wI];
    n_live += sm.nlive[wI];
    if (first_dead < 0) first_dead = sm.fdead[wI];
  }
  if (live) {
    const int s = before + __popc(live_mask & ((1u << lane) - 1u));
    P.lay[s] = tid; P.gat[s] = my_slot; P.mlp_lay[s] = lp;
  }
  if (tid == first_dead) {          // the ghost slot stands for every dead cluster
    P.lay[n_live] = tid; P.gat[n_live] = my_slot; P.mlp_lay[n_live] = lp;
  }
  const int nl = n_live + (first_dead >= 0 ? 1 : 0);
  for (int s = nl + tid; s < 8 * ((nl + 7) / 8); s += NT) { P.lay[s] = -1; P.gat[s] = 0; P.mlp_lay[s] = 0.0; }
  // publish: everything this update wrote is visible before the counter moves
  SU_TICK(4);
  __threadfence();
  vl_bar_sync<BAR>(NT);
  if (tid == 0) *reinterpret_cast<volatile int*>(&S->pub) = R.n_updates + 1;
  SU_TICK(5);
  return 0;
}
