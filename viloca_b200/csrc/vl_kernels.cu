// CAVI kernels for sm_100a.  One iteration (update) of a (window, restart) problem = its
// haplotype tiles, then its read tiles; vl_chunk_kernel runs a whole chunk of updates of all
// scheduled restarts in one launch (blocks wait on per-restart counters), vl_haplo_kernel and
// vl_cluster_kernel run the two phases of one update as separate launches (same device code):
//   vl_haplo_body      update_mean_haplo   (uqs update_eqs.py:73-101, lep update_eqs.py:115-159)
//                      + the sums update_a_and_b / elbo_haplo need (update_eqs.py:103-112,
//                      elbo_eqs.py:65-78)
//   vl_cluster_body    update_mean_cluster (uqs update_eqs.py:55-71, lep update_eqs.py:90-112)
//                      + the sums update_alpha / elbo_data / elbo_cluster / update_c_and_d need;
//                      the CTA of a restart that finishes last then does the scalar update
//                      (vl_cavi.cuh): alpha, mean_log_pi, a, b, (c, d), the ELBO, the loop state
//                      machine of run_cavi (cavi.py:120-169) and the next live-cluster layout
//
// Both contractions run on the FP64 tensor pipe (DMMA m8n8k4) over the dense-column operand
// Dm (vl_device.cuh), whose tiles and the other operand (w z or the haplotype complement
// table) are staged in shared memory by TMA bulk copies behind mbarriers; the remaining
// entries are added from the sparse lists; softmax / ELBO reductions are fused into the epilogues.
#ifdef VL_RS_PROFILE
#define VL_SU_PROFILE
__device__ unsigned long long vl_su_cycles[8];     // vl_scalar_update phases of restart 0 (vl_cavi.cuh)
#endif
#include "vl_cavi.cuh"
#include <cstdio>
#include <cstdlib>

#include "vl_launch.h"
#ifndef VL_ATTR_KB
#define VL_ATTR_KB 220
#endif
#include "../../include/viloca_b200.h"

// Per-phase cycle counters of restart 0's tile 0 (development aid, -DVL_RS_PROFILE):
// haplotype tile: 0 prologue, 1 operand loop, 2 warp reduction, 3 sparse, 4 softmax epilogue, 5 sums;
// read tile: 8 prologue, 9 operand loop, 10 sparse, 11 softmax epilogue, 12 records + ticket,
// 13 scalar update (whichever tile is last); chunk kernel: 6 wait published (haplotype tile),
// 7 signal, 14 wait published (read tile), 15 wait haplotype tiles; 16 / 17 samples.
#ifdef VL_RS_PROFILE
__device__ unsigned long long vl_gk_cycles[20];
#define GK_DECL(cond) const bool gk_on_ = (cond) && threadIdx.x == 0; long long gk_t_ = clock64()
#define GK_TICK(slot)                                                                   \
  do {                                                                                  \
    if (gk_on_) {                                                                       \
      const long long n_ = clock64();                                                   \
      atomicAdd(&vl_gk_cycles[slot], (unsigned long long)(n_ - gk_t_));                 \
      gk_t_ = clock64();                                                                \
    }                                                                                   \
  } while (0)
#else
#define GK_DECL(cond) do { } while (0)
#define GK_TICK(slot) do { } while (0)
#endif

// CTAs per SM each kernel class is compiled for (bounds the registers per thread): the work of a
// narrow layout is latency bound, so residency matters more than registers there
#ifndef VL_C0_CTAS
#define VL_C0_CTAS 6       // resident CTAs per SM the narrowest class is compiled for
#endif
#ifndef VL_C0_STAGES
#define VL_C0_STAGES 3     // its operand stages
#endif
constexpr int vl_min_ctas(int ktm) { return ktm <= 2 ? VL_C0_CTAS : (ktm <= 4 ? 5 : (ktm <= 8 ? 3 : 2)); }
__host__ __device__ constexpr int vl_zstages(int ktm) { return ktm <= 2 ? VL_C0_STAGES : VL_ZSTAGES; }
// operand stages of the haplotype tile: the widest class double-buffers so that two CTAs fit one SM
// (three stages of 32 reads x 132 doubles would leave room for one: tools/occupancy.py)
__host__ __device__ constexpr int vl_hstages(int ktm) { return ktm > 8 ? 2 : (ktm <= 2 ? VL_C0_STAGES : VL_STAGES); }
// doubles of dynamic shared memory a haplotype tile needs: its operand stages, or the post-loop tables
// T [16][K8] + Cm [16][5][K8] that reuse them, whichever is larger
__host__ __device__ constexpr int vl_haplo_doubles(int ktm) {
  const int stages = vl_hstages(ktm) * (VL_H_NC * VL_PT + VL_H_NC * vl_ks(ktm)), tables = VL_PT * 6 * 8 * ktm;
  return stages > tables ? stages : tables;
}

namespace {

// Poll (thread 0, bounded) until the restart has published `target` completed updates; false if
// it left its loop instead.
__device__ __forceinline__ bool vl_wait_published(const VlState* S, int target) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    const volatile int* pub = &S->pub;
    const volatile int* act = &S->active;
    int r = 1;
    for (unsigned spin = 0; *pub < target; ++spin) {
      if (*act == 0) { r = 0; break; }
      __nanosleep(spin < 4096 ? 40 : 1000);
      if (spin > (1u << 24)) __trap();       // bounded (~20 s): a lost dependency is an error, not a hang
    }
    if (*act == 0) r = 0;
    ok = r;
    __threadfence();
  }
  __syncthreads();
  const int r = ok;
  __syncthreads();
  return r != 0;
}

// A column tile whose reads are shared by n_split CTAs: park this CTA's partial sums T; the CTA that
// arrives last adds all parts in split order (a fixed order, whoever arrives last) into its T and
// carries on (returns true), the others are done.  Kept out of line: windows below 1024 padded
// reads never come here and their code should not pay for it.
__device__ __noinline__ bool vl_split_combine(const VlProblem& P, double* T, int ty, int split, int n_split, int count) {
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const size_t part_stride = (size_t)VL_PT * 8 * P.KT;
  double* parts = P.tsplit + (size_t)ty * n_split * part_stride;
  __syncthreads();
  for (int idx = tid; idx < count; idx += VL_THREADS) parts[split * part_stride + idx] = T[idx];
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&P.hsplit_count[ty], 1) == n_split - 1) ? 1 : 0;
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  for (int idx = tid; idx < count; idx += VL_THREADS) {
    double v = parts[idx];
    for (int s2 = 1; s2 < n_split; ++s2) v += parts[s2 * part_stride + idx];
    T[idx] = v;
  }
  if (tid == 0) P.hsplit_count[ty] = 0;          // the next update's CTAs of this tile come after our signal
  return true;
}

// =========================================================================================
// update_mean_haplo.  Per (position l, base b, slot s) the logit is
//   (l, b) a dense column j:  T[j,s]    = sum_n w_n z[n,s] Dm[n,j]              (DMMA, M = column)
//   otherwise:                Cm[l,b,s] = sum_{sparse entries (n,l,b)} (w_n d) z[n,s]
// plus the reference term; the part of the reference's logits that does not depend on b is
// dropped (it cancels in the softmax).  CTA = 16 columns (whole positions) x all slots; warp
// (lg, rh) = 8 columns x one half of the reads of every stage; the halves are added in fixed
// order.
// =========================================================================================
// SPLIT: the variant for large windows whose tiles share their reads among several CTAs; windows below
// 1024 padded reads run the SPLIT = false instantiation, which is the code they always had.
template <int KTM, int MODE, bool SPLIT>
__device__ __forceinline__ bool vl_haplo_body(const VlProblem* __restrict__ probs, const int2 tm, const int stage_doubles,
                                              const bool early, const int pub_want) {
  constexpr int NB = (8 * KTM + 31) / 32;        // 32-slot groups (lanes = slots in the sparse part)
  constexpr int KSM = vl_ks(KTM);
  constexpr int DM_STAGE = VL_H_NC * VL_PT;            // doubles
  extern __shared__ __align__(128) unsigned char vl_smem[];
  double* stg = reinterpret_cast<double*>(vl_smem);
  __shared__ uint64_t bars[VL_MAX_STAGES];
  __shared__ double red[VL_THREADS / 32][4];

  GK_DECL(tm.x == 0 && (tm.y & 0xFFFFF) == 0);     // (all read parts of column tile 0: the loop phases are summed over them)
  const VlProblem& P = probs[tm.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  // Large windows: n_hsplit CTAs share the reads of a column tile;
  // the last one to arrive adds the partial sums in split order and finishes the tile.
  // (the split travels in the tile id -- vl_htile_id -- so that no tile waits for a load to learn its tile)
  const int ty = SPLIT ? (tm.y & 0xFFFFF) : tm.y, split = SPLIT ? ((tm.y >> 20) & 0xF) : 0;
  const int n_split = SPLIT ? max(1, (tm.y >> 24) & 0xF) : 1;
  // Inside the scheduler kernel (early) the tile is handed out as soon as the read tiles of the
  // previous update have written w z: the operand loop needs that update's layout width only
  // (ktring) and overlaps the scalar update; the state is read after it has been published.
  // This tile's static maps go to shared memory up front (one memory latency, overlapped).
  __shared__ VlState sS;
  __shared__ int s_tpos[VL_PT], s_collb[VL_PT], s_poscol[VL_PT][VL_B], s_ep[VL_PT][VL_B + 1];
  __shared__ unsigned s_refb[VL_PT];
  __shared__ int s_kt_old;
  if (tid < VL_PT) {
    s_tpos[tid] = P.tile_pos[(size_t)ty * VL_PT + tid];
    s_collb[tid] = P.col_lb[(size_t)ty * VL_PT + tid];
  }
  if (early) {
    if (tid == 0) {
      s_kt_old = *reinterpret_cast<const volatile int*>(&P.st->ktring[(pub_want + 1) & 1]);
      __threadfence();
    }
  } else {
    vl_load_state(&sS, P.st, tid);
  }
  __syncthreads();
  if (!early && tid == 0) s_kt_old = (sS.active != 0) ? sS.ktl_z : -1;
  if (!early) __syncthreads();
  const int ktz = s_kt_old;                // tiles of the layout w z is stored in
  if (ktz < 0 || ktz > KTM) return false;
  if (early) asm volatile("fence.proxy.async;\n" ::: "memory");   // bulk copies below read what other CTAs wrote
  const int Npad = P.Npad;
  const int KSz = vl_ks(ktz), K8z = 8 * ktz;
  const int nchunks_all = Npad / VL_H_NC;
  const int per_split = (nchunks_all + n_split - 1) / n_split;
  const int c_first = min(nchunks_all, split * per_split);            // this CTA's chunks of 32 reads
  const int nchunks = min(nchunks_all, c_first + per_split) - c_first;
  // A stage = `grp` consecutive chunks (contiguous in dm and in w z: two bulk copies, one barrier phase).
  // With shared memory to spare (narrow layout, or the deep launch of a chunk with few tiles in flight)
  // the stages grow: an update of a lone restart is a chain of latencies, and every stage costs one
  // issue by one thread and one CTA-wide barrier.  The rows are visited in the same order whatever grp.
  const int chunk_doubles = DM_STAGE + VL_H_NC * vl_ks(ktz);
  constexpr int OWN_DOUBLES = vl_haplo_doubles(KTM);
  const int fit = max(stage_doubles, OWN_DOUBLES) / chunk_doubles;     // chunks the shared memory holds
#ifdef VL_NO_GROUP
  const int grp = 1;
#else
  const int grp = (fit >= nchunks) ? max(1, (nchunks + 2) / 3) : max(1, min(8, fit / 3));
#endif
  const int st_stride = grp * chunk_doubles;
  const int nstg = max(1, min(VL_MAX_STAGES, fit / grp));
  const int nstages = (nchunks + grp - 1) / grp;
  const double* __restrict__ dmg = P.dm + (size_t)ty * Npad * VL_PT + (size_t)c_first * VL_H_NC * VL_PT;
  const double* wzg = P.wz;     // (mutable state: no __restrict__, these must stay coherent loads)
  const double* wzs = wzg + (size_t)c_first * VL_H_NC * vl_ks(ktz);   // first read of this CTA's range
  const int* gat = P.gat;
  const uint32_t z_bytes = (uint32_t)(VL_H_NC * KSz * sizeof(double));       // per chunk

  if (tid == 0) {
    for (int s = 0; s < nstg; ++s) vl_mbar_init(&bars[s], 1);
    vl_fence_mbar_init();
  }
  if (tid < VL_PT * (VL_B + 1)) {
    const int pos = tid / (VL_B + 1), j = tid - pos * (VL_B + 1);
    const int l = s_tpos[pos];
    s_ep[pos][j] = (l >= 0) ? P.sp_ptr[l * VL_B + j] : 0;
    if (j < VL_B) s_poscol[pos][j] = (l >= 0) ? P.poscol[l * VL_B + j] : -1;
    if (j == VL_B) s_refb[pos] = (l >= 0) ? P.ref[l] : 255u;
  }
  __syncthreads();
  // stage c: chunks [c grp, c grp + n); shared memory: n dm chunks, then (at grp * DM_STAGE) their w z rows
  auto issue = [&](int c) {
    const int sn = c % nstg;
    const int n = min(grp, nchunks - c * grp);
    double* dst = stg + (size_t)sn * st_stride;
    vl_mbar_expect_tx(&bars[sn], (uint32_t)n * ((uint32_t)(DM_STAGE * sizeof(double)) + z_bytes));
    vl_tma_load_1d(dst, dmg + (size_t)c * grp * DM_STAGE, (uint32_t)n * DM_STAGE * sizeof(double), &bars[sn]);
    vl_tma_load_1d(dst + grp * DM_STAGE, wzs + (size_t)c * grp * VL_H_NC * KSz, (uint32_t)n * z_bytes, &bars[sn]);
  };
  if (tid == 0)
    for (int c = 0; c < nstg - 1 && c < nstages; ++c) issue(c);

  // warp = all 16 columns (two m-tiles) x 8 of the 32 reads of every stage: one w z fragment
  // feeds two DMMAs
  double acc[2][KTM][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int i = 0; i < KTM; ++i) { acc[mt][i][0] = 0.0; acc[mt][i][1] = 0.0; }

  const int acol0 = (g + 4 * q) & 15, acol1 = (8 + g + 4 * q) & 15;   // vl_dm_col(row, g / 8 + g), row % 4 == q
  GK_TICK(0);
  vl_dispatch_tiles<KTM>(ktz, [&](auto ktc) {
    constexpr int KT = decltype(ktc)::value;
    auto chunk = [&](const double* dt, const double* zt) {     // 32 reads: rows warp * 8 + 4 s + q
#pragma unroll
      for (int s = 0; s < VL_H_NC / 16; ++s) {
        const int row = warp * (VL_H_NC / 4) + 4 * s + q;
        const double a0 = dt[row * VL_PT + acol0];
        const double a1 = dt[row * VL_PT + acol1];
        const double* zr = zt + row * KSz + g;
#pragma unroll
        for (int i = 0; i < KT; ++i) {
          const double bf = zr[8 * i];
          vl_dmma(acc[0][i], a0, bf);
          vl_dmma(acc[1][i], a1, bf);
        }
      }
    };
    if (grp == 1) {                    // (the loop of the busy phases, kept free of the grouping arithmetic)
      for (int c = 0; c < nchunks; ++c) {
        const int st = c % nstg;
        if (tid == 0 && c + nstg - 1 < nchunks) issue(c + nstg - 1);
        vl_mbar_wait(&bars[st], (c / nstg) & 1);
        const double* dt = stg + (size_t)st * st_stride;
        chunk(dt, dt + DM_STAGE);
        __syncthreads();
      }
    } else {
      for (int c = 0; c < nstages; ++c) {
        const int st = c % nstg;
        if (tid == 0 && c + nstg - 1 < nstages) issue(c + nstg - 1);
        vl_mbar_wait(&bars[st], (c / nstg) & 1);
        const int n = min(grp, nchunks - c * grp);
        const double* dt = stg + (size_t)st * st_stride;
        const double* zt = dt + grp * DM_STAGE;
        for (int u = 0; u < n; ++u) chunk(dt + u * DM_STAGE, zt + u * VL_H_NC * KSz);
        __syncthreads();
      }
    }
  });

  // (every thread is past its last wait: the barriers are given back so that the next tile this
  // CTA runs -- the scheduler kernel's workers are persistent -- may initialise them again)
  if (tid == 0)
    for (int s = 0; s < nstg; ++s) vl_mbar_inval(&bars[s]);
  // ---- the stage buffers are free now: T [16][K8] and Cm [16][5][K8] live in them; the four
  // warps add their partial sums over the reads into T one after the other (fixed order) ----
  GK_TICK(1);
  // T [16][8 ktz] in the slot order of the stored w z
  double* T = stg;
  auto combine_warps = [&]() {
    for (int turn = 0; turn < VL_THREADS / 32; ++turn) {
      if (warp == turn) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int i = 0; i < KTM; ++i)
            if (i < ktz) {
              double2* tp = reinterpret_cast<double2*>(T + (mt * 8 + g) * K8z + 8 * i + 2 * q);
              double2 o = make_double2(0.0, 0.0);
              if (turn > 0) o = *tp;
              *tp = make_double2(o.x + acc[mt][i][0], o.y + acc[mt][i][1]);
            }
      }
      if (turn + 1 < VL_THREADS / 32) __syncthreads();
    }
  };
  if constexpr (SPLIT) {
    combine_warps();
    if (n_split > 1 && !vl_split_combine(P, T, ty, split, n_split, VL_PT * K8z)) return false;
  }
  GK_TICK(2);
  // the state of this update: published by the scalar update of the previous one
  if (early) {
    if (!vl_wait_published(P.st, pub_want)) return false;
    vl_load_state(&sS, P.st, tid);
    __syncthreads();
  }
  GK_TICK(6);
  const VlState* S = &sS;
  const int ktl = S->ktl;
  if (S->active == 0 || ktl > KTM || S->ktl_z != ktz) return false;
  const int KSn = vl_ks(ktl), K8 = 8 * ktl;
  const int nlay = S->nlay;
  // Cm [16][5][K8] in the current slot order, behind T
  double* Cm = stg + VL_PT * K8z;
  if constexpr (!SPLIT) combine_warps();
  GK_TICK(2);
  // sparse entries of the positions whose columns lie in this tile: warp = 4 of them, lanes = slots.
  // The warp's entries are four contiguous ranges of the (position, base)-sorted list; they are fetched 32 at
  // a time (lane l = entry l: read index, weight, target), then handed round by shuffles eight at a time so
  // that eight rows of w z are in flight instead of one dependent load after the other (these loops were
  // 10 % of a lone restart's update).  Every (position, base) sum adds its entries in list order, as before.
  {
    const int* __restrict__ sp_n = P.sp_n;
    const double* __restrict__ sp_d = P.sp_d;
    constexpr int U = 8;
    int gl[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int s = lane + 32 * j;
      gl[j] = (s < nlay) ? gat[s] : 0;
    }
    int r0[4], rn[4];
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) {
      const int pos = warp * 4 + pp;
      const bool valid = s_tpos[pos] >= 0;
      r0[pp] = valid ? s_ep[pos][0] : 0;
      rn[pp] = valid ? s_ep[pos][VL_B] - r0[pp] : 0;
      if (valid)                                          // a (position, base) pair without entries keeps this zero
        for (int b = 0; b < VL_B; ++b)
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const int sl = lane + 32 * j;
            if (sl < K8) Cm[(pos * VL_B + b) * K8 + sl] = 0.0;
          }
    }
    const int total = rn[0] + rn[1] + rn[2] + rn[3];
    double av[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) av[j] = 0.0;
    int cur = -1;                                         // pos * VL_B + b of the running sums
    auto flush = [&]() {
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int sl = lane + 32 * j;
        if (sl < K8) Cm[cur * K8 + sl] = av[j];
      }
    };
    for (int base = 0; base < total; base += 32) {
      const int cnt = min(32, total - base);
      int my_n = 0, my_seg = 0;
      double my_d = 0.0;
      if (lane < cnt) {
        int i = base + lane, pp = 0;
        if (i >= rn[0]) { i -= rn[0]; pp = 1; if (i >= rn[1]) { i -= rn[1]; pp = 2; if (i >= rn[2]) { i -= rn[2]; pp = 3; } } }
        const int pos = warp * 4 + pp;
        const int e = (pp == 0 ? r0[0] : pp == 1 ? r0[1] : pp == 2 ? r0[2] : r0[3]) + i;
        my_n = sp_n[e];
        my_d = sp_d[e];
        int bb = 0;
        while (bb < VL_B - 1 && e >= s_ep[pos][bb + 1]) ++bb;
        my_seg = pos * VL_B + bb;
      }
      for (int k0 = 0; k0 < cnt; k0 += U) {
        double v[U][NB], dk[U];
        int sg[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int kk = min(k0 + u, 31);
          const int nk = __shfl_sync(0xffffffffu, my_n, kk);
          dk[u] = __shfl_sync(0xffffffffu, my_d, kk);
          sg[u] = __shfl_sync(0xffffffffu, my_seg, kk);
          const double* zr = wzg + (size_t)nk * KSz;      // (entries past cnt read a valid row and are not used)
#pragma unroll
          for (int j = 0; j < NB; ++j) v[u][j] = zr[gl[j]];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (k0 + u < cnt) {
            if (sg[u] != cur) {
              if (cur >= 0) flush();
              cur = sg[u];
#pragma unroll
              for (int j = 0; j < NB; ++j) av[j] = 0.0;
            }
#pragma unroll
            for (int j = 0; j < NB; ++j) av[j] += dk[u] * v[u][j];
          }
        }
      }
    }
    if (cur >= 0) flush();
  }
  __syncthreads();

  GK_TICK(3);
  // ---- epilogue: + ref_part, softmax over b, the three sums, store h, g, gm ----
  const double mlg0 = S->mlg0, refoff = S->mlg1 - VL_LOG4;
  double cscale = 1.0;
  if (MODE == VL_MODE_LEP) cscale = S->mlt0 - (S->mlt1 - VL_LOG4);
  const int ghost = S->ghost;
  const double gmult = S->ghost_mult;
  double* __restrict__ gmt = P.gm + (size_t)ty * VL_PT * KSn;
  double s_ref = 0.0, s_nref = 0.0, s_ent = 0.0;
  for (int idx = tid; idx < VL_PT * K8; idx += VL_THREADS) {
    const int pos = idx / K8, slot = idx - pos * K8;
    if (s_collb[pos] < 0) gmt[pos * KSn + slot] = 0.0;     // padding column of the operand
    const int l = s_tpos[pos];
    if (l < 0) continue;
    const unsigned rc = s_refb[pos];
    double* hp = P.h + (size_t)l * VL_B * KSn + slot;
    double* gp = P.g + (size_t)l * VL_B * KSn + slot;
    int lc[VL_B];                                          // column of (l, b) within this tile, or -1
#pragma unroll
    for (int b = 0; b < VL_B; ++b) {
      const int j = s_poscol[pos][b];
      lc[b] = (j < 0) ? -1 : j - ty * VL_PT;
    }
    if (slot < nlay) {
      const int zslot = gat[slot];                         // column of the stored w z this slot continues
      double lgt[VL_B];
#pragma unroll
      for (int b = 0; b < VL_B; ++b) {
        const double cb = (lc[b] >= 0) ? T[lc[b] * K8z + zslot] : Cm[(pos * VL_B + b) * K8 + slot];
        lgt[b] = ((MODE == VL_MODE_LEP) ? cscale * cb : cb) + ((rc == (unsigned)b) ? mlg0 : refoff);
      }
      double m = lgt[0];
#pragma unroll
      for (int b = 1; b < VL_B; ++b) m = fmax(m, lgt[b]);
      double e[VL_B], xs[VL_B], sum = 0.0;
#pragma unroll
      for (int b = 0; b < VL_B; ++b) xs[b] = lgt[b] - m;
      vl_exp_nonpos<VL_B>(xs, e);
#pragma unroll
      for (int b = 0; b < VL_B; ++b) sum += e[b];
      const double ls = log(sum);
      const double rsum = 1.0 / sum;
      const double mu = (slot == ghost) ? gmult : 1.0;
      double r0 = 0.0, r1 = 0.0, r2 = 0.0;
#pragma unroll
      for (int b = 0; b < VL_B; ++b) {
        const double hb = e[b] * rsum;
        double oth = 0.0;
#pragma unroll
        for (int b2 = 0; b2 < VL_B; ++b2) if (b2 != b) oth += e[b2];
        const double gb = oth * rsum;
        if (rc == (unsigned)b) r0 += hb; else r1 += hb;
        if (hb > 0.0) r2 += hb * ((lgt[b] - m) - ls);
        hp[b * KSn] = hb;
        gp[b * KSn] = gb;
        if (lc[b] >= 0) gmt[lc[b] * KSn + slot] = gb;
      }
      s_ref += mu * r0; s_nref += mu * r1; s_ent += mu * r2;
    } else {
#pragma unroll
      for (int b = 0; b < VL_B; ++b) {
        hp[b * KSn] = 0.0; gp[b * KSn] = 0.0;
        if (lc[b] >= 0) gmt[lc[b] * KSn + slot] = 0.0;
      }
    }
  }
  GK_TICK(4);
  s_ref = warp_sum(s_ref); s_nref = warp_sum(s_nref); s_ent = warp_sum(s_ent);
  if (lane == 0) { red[warp][0] = s_ref; red[warp][1] = s_nref; red[warp][2] = s_ent; }
  __syncthreads();
  if (tid < 3) {
    double v = 0.0;
#pragma unroll
    for (int wI = 0; wI < VL_THREADS / 32; ++wI) v += red[wI][tid];
    P.hpart[(size_t)ty * VL_HP_STRIDE + tid] = v;
  }
  GK_TICK(5);
#ifdef VL_RS_PROFILE
  if (gk_on_) atomicAdd(&vl_gk_cycles[16], 1ull);
#endif
  return true;
}

template <int KTM, int MODE, bool SPLIT>
__global__ void __launch_bounds__(VL_THREADS, vl_min_ctas(KTM))
vl_haplo_kernel(const VlProblem* __restrict__ probs, const int2* __restrict__ tiles) {
  vl_haplo_body<KTM, MODE, SPLIT>(probs, tiles[blockIdx.x], 0, false, 0);
}

// =========================================================================================
// update_mean_cluster.  Y[n,s] = sum_l Dm[n,l] gm[l,s] (dense, DMMA, M = read) + the minority
// entries of read n; uqs logits = ltsum[n] - Y + mean_log_pi, softmax over the slots with the
// ghost slot counted ghost_mult times.  CTA = 64 reads x all slots; warp = 16 reads.
// =========================================================================================
// vl_cluster_tile: one tile up to its per-tile record; returns false if the restart is not running in this
// kernel class.  vl_cluster_close: what the tile that finishes last does for its restart.  sS receives the
// restart's state as it was before the update.
// MT = m-tiles of 8 reads per warp: 2 (CTA = 64 reads), or 1 (CTA = 32 reads) for the windows whose read
// tiles would otherwise occupy a fraction of the SMs (VlProblem::zt_reads, a function of the window alone).
template <int KTM, int MODE, int MT>
__device__ __forceinline__ bool vl_cluster_tile_mt(const VlProblem* __restrict__ probs, const int2 tm,
                                                   const int stage_doubles, VlState& sS) {
  constexpr int KSM = vl_ks(KTM), K8M = 8 * KTM;
  constexpr int ZT = 32 * MT;                           // reads of this tile
  constexpr int DM_STAGE = ZT * VL_PT;                  // doubles
  constexpr int ST_STRIDE = VL_ZT_READS * VL_PT + VL_PT * KSM;    // dm | gm as the shared memory is sized (64 reads)
  extern __shared__ __align__(128) unsigned char vl_smem[];
  double* stg = reinterpret_cast<double*>(vl_smem);
  double* red_cs = stg + max(stage_doubles, vl_zstages(KTM) * ST_STRIDE);   // [4][K8M], after the stages
  __shared__ uint64_t bars[VL_MAX_STAGES];
  __shared__ double red_sc[VL_THREADS / 32][4];

  GK_DECL(tm.x == 0 && tm.y == 0);
  const VlProblem& P = probs[tm.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  // state and mean_log_pi of the slots in shared memory: one latency, off the epilogue's path
  __shared__ double s_mlp[VL_MAX_CLUSTERS];
  vl_load_state(&sS, P.st, tid);
  s_mlp[tid] = (tid < 8 * P.KT) ? __ldcg(P.mlp_lay + tid) : 0.0;
  __syncthreads();
  const VlState* S = &sS;
  const int ktl = S->ktl;
  if (S->active == 0 || ktl > KTM || S->ktl_z > KTM) return false;

  const int Npad = P.Npad;
  const int KS = vl_ks(ktl), K8 = 8 * ktl;
  const int nst = P.NCs / VL_PT;
  const int n0 = tm.y * ZT;
  const double* __restrict__ dmg = P.dm + (size_t)n0 * VL_PT;
  const double* gmg = P.gm;
  const uint32_t gm_bytes = (uint32_t)(VL_PT * KS * sizeof(double));          // per column tile
  // a stage = `grp` consecutive column tiles (their gm rows are contiguous: one copy; one dm copy per
  // tile; one barrier phase): larger stages when shared memory is to spare, see vl_haplo_body
  const int tile_doubles = DM_STAGE + VL_PT * KS;
  const int fit = max(stage_doubles, vl_zstages(KTM) * ST_STRIDE) / tile_doubles;
#ifdef VL_NO_GROUP
  const int grp = 1;
#else
  const int grp = (fit >= nst) ? max(1, (nst + 2) / 3) : max(1, min(8, fit / 3));
#endif
  const int st_stride = grp * tile_doubles;
  const int nstg = max(1, min(VL_MAX_STAGES, fit / grp));
  const int nstages = (nst + grp - 1) / grp;

  if (tid == 0) {
    for (int s = 0; s < nstg; ++s) vl_mbar_init(&bars[s], 1);
    vl_fence_mbar_init();
  }
  __syncthreads();
  // stage c: column tiles [c grp, c grp + n); shared memory: n dm tiles, then (at grp * DM_STAGE) their gm rows
  auto issue = [&](int c) {
    const int sn = c % nstg;
    const int n = min(grp, nst - c * grp);
    double* dst = stg + (size_t)sn * st_stride;
    vl_mbar_expect_tx(&bars[sn], (uint32_t)n * ((uint32_t)(DM_STAGE * sizeof(double)) + gm_bytes));
    for (int u = 0; u < n; ++u)
      vl_tma_load_1d(dst + u * DM_STAGE, dmg + (size_t)(c * grp + u) * Npad * VL_PT, DM_STAGE * sizeof(double), &bars[sn]);
    vl_tma_load_1d(dst + grp * DM_STAGE, gmg + (size_t)c * grp * VL_PT * KS, (uint32_t)n * gm_bytes, &bars[sn]);
  };
  if (tid == 0)
    for (int c = 0; c < nstg - 1 && c < nstages; ++c) issue(c);

  double acc[MT][KTM][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int i = 0; i < KTM; ++i) { acc[mt][i][0] = 0.0; acc[mt][i][1] = 0.0; }

  const int r0 = warp * (8 * MT) + g;              // rows r0 (and r0 + 8) of the tile; r0 % 4 == g % 4
  // per-read scalars: requested now, needed after the operand loop
  double wn_r[MT], mc_r[MT], lt_r[MT];
  int se0[MT], se1[MT];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int n = n0 + r0 + 8 * mt;
    wn_r[mt] = P.w[n];
    mc_r[mt] = P.mcount[n];
    lt_r[mt] = (MODE == VL_MODE_UQS) ? P.ltsum[n] : 0.0;
    se0[mt] = P.sr_ptr[n];
    se1[mt] = P.sr_ptr[n + 1];
  }
  GK_TICK(8);
  vl_dispatch_tiles<KTM>(ktl, [&](auto ktc) {
    constexpr int KT = decltype(ktc)::value;
    auto tile = [&](const double* dt, const double* gt) {      // one column tile: 16 columns x this warp's 16 reads
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int col = (4 * s + q + 4 * (g & 3)) & 15;
        const double a0 = dt[r0 * VL_PT + col];
        const double a1 = (MT > 1) ? dt[(r0 + 8) * VL_PT + col] : 0.0;
        const double* gr = gt + (4 * s + q) * KS;
#pragma unroll
        for (int i = 0; i < KT; ++i) {
          const double bf = gr[8 * i];
          vl_dmma(acc[0][i], a0, bf);
          if (MT > 1) vl_dmma(acc[MT - 1][i], a1, bf);
        }
      }
    };
    if (grp == 1) {                    // (the loop of the busy phases, kept free of the grouping arithmetic)
      for (int c = 0; c < nst; ++c) {
        const int st = c % nstg;
        if (tid == 0 && c + nstg - 1 < nst) issue(c + nstg - 1);
        vl_mbar_wait(&bars[st], (c / nstg) & 1);
        const double* dt = stg + (size_t)st * st_stride;
        tile(dt, dt + DM_STAGE + g);
        __syncthreads();
      }
    } else {
      for (int c = 0; c < nstages; ++c) {
        const int st = c % nstg;
        if (tid == 0 && c + nstg - 1 < nstages) issue(c + nstg - 1);
        vl_mbar_wait(&bars[st], (c / nstg) & 1);
        const int n = min(grp, nst - c * grp);
        const double* dt = stg + (size_t)st * st_stride;
        const double* gt = dt + grp * DM_STAGE + g;
        for (int u = 0; u < n; ++u) tile(dt + u * DM_STAGE, gt + u * VL_PT * KS);
        __syncthreads();
      }
    }
  });

  if (tid == 0)
    for (int s = 0; s < nstg; ++s) vl_mbar_inval(&bars[s]);
  GK_TICK(9);
  // ---- sparse entries of this thread's two reads ----
  {
    const int* __restrict__ sr_idx = P.sr_idx;
    const double* __restrict__ sr_d = P.sr_d;
    const double* gtab = P.g;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      for (int e = se0[mt]; e < se1[mt]; ++e) {
        const double d = sr_d[e];
        const double* gp = gtab + (size_t)sr_idx[e] * KS + 2 * q;
#pragma unroll
        for (int i = 0; i < KTM; ++i)
          if (i < ktl) {
            const double2 gv = *reinterpret_cast<const double2*>(gp + 8 * i);
            acc[mt][i][0] += d * gv.x;
            acc[mt][i][1] += d * gv.y;
          }
      }
    }
  }

  GK_TICK(10);
  // ---- epilogue: logits, softmax over the slots, the reductions, store mean_cluster ----
  const double* mlpl = s_mlp;
  const int N = P.N, nlay = S->nlay, ghost = S->ghost;
  const double gmult = S->ghost_mult;
  double b1 = 0.0, b2 = 0.0;
  const double Lf = (double)P.L;
  if (MODE == VL_MODE_LEP) { b1 = S->mlt0; b2 = S->mlt1 - VL_LOG4; }
  double s_data = 0.0, s_ent = 0.0, s_c = 0.0, s_d = 0.0;
  double cs[KTM][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int n = n0 + r0 + 8 * mt;
    const bool valid = n < N;
    const double wn = wn_r[mt];
    const double mc = mc_r[mt];
    const double base = (MODE == VL_MODE_UQS) ? lt_r[mt] : mc;
    double t[KTM][2];
    double m = neg_inf();
#pragma unroll
    for (int i = 0; i < KTM; ++i)
      if (i < ktl) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int s = 8 * i + 2 * q + j;
          // uqs: X = sum_{l,b} E h;  lep: S = sum_{l,b} r h  (update_eqs.py:57 / lep :98-99)
          const double x = base - acc[mt][i][j];
          acc[mt][i][j] = x;
          double tv = neg_inf();
          if (s < nlay) tv = ((MODE == VL_MODE_UQS) ? x : (b1 * x + b2 * (Lf - x))) + mlpl[s];
          t[i][j] = tv;
          m = fmax(m, tv);
        }
      }
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
    double sum = 0.0, u = 0.0, v = 0.0;
#pragma unroll
    for (int i = 0; i < KTM; ++i)
      if (i < ktl) {
        const double xs[2] = {t[i][0] - m, t[i][1] - m};
        double ev[2];
        vl_exp_nonpos<2>(xs, ev);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int s = 8 * i + 2 * q + j;
          double e = 0.0;
          if (s < nlay) {
            e = ev[j];
            const double em = (s == ghost) ? gmult * e : e;
            if (e > 0.0) u += em * xs[j];
            v += em * acc[mt][i][j];
            sum += em;
          }
          t[i][j] = e;
        }
      }
    const double psum = sum;   // this thread's share of the row sum
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    const double rsum = 1.0 / sum;
    if (valid) {
      // sum_k z log z = (sum_k e_k (t_k - m)) / s - log s   (0 log 0 := 0)
      s_ent += u * rsum - ((q == 0) ? log(sum) : 0.0);
      if (MODE == VL_MODE_UQS) {
        s_data += wn * (v * rsum);
      } else {
        s_c += wn * (v * rsum);
        s_d += wn * ((mc * psum - v) * rsum);
      }
    }
    double* zrow = P.z + (size_t)n * KS + 2 * q;
    double* wzrow = P.wz + (size_t)n * KS + 2 * q;
#pragma unroll
    for (int i = 0; i < KTM; ++i)
      if (i < ktl) {
        const double z0 = valid ? t[i][0] * rsum : 0.0;
        const double z1 = valid ? t[i][1] * rsum : 0.0;
        const double wz0 = wn * z0, wz1 = wn * z1;
        *reinterpret_cast<double2*>(zrow + 8 * i) = make_double2(z0, z1);
        *reinterpret_cast<double2*>(wzrow + 8 * i) = make_double2(wz0, wz1);
        if (mt == 0) { cs[i][0] = wz0; cs[i][1] = wz1; }
        else { cs[i][0] += wz0; cs[i][1] += wz1; }
      }
  }
  GK_TICK(11);
#pragma unroll
  for (int i = 0; i < KTM; ++i)
    if (i < ktl) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        double v = cs[i][j];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0) red_cs[warp * K8M + 8 * i + 2 * q + j] = v;
      }
    }
  s_data = warp_sum(s_data); s_ent = warp_sum(s_ent); s_c = warp_sum(s_c); s_d = warp_sum(s_d);
  if (lane == 0) {
    red_sc[warp][0] = s_data; red_sc[warp][1] = s_ent; red_sc[warp][2] = s_c; red_sc[warp][3] = s_d;
  }
  __syncthreads();
  const int K8rec = 8 * P.KT, ZS = K8rec + VL_ZP_EXTRA;
  double* zp = P.zpart + (size_t)tm.y * ZS;
  for (int k = tid; k < K8 + 4; k += VL_THREADS) {
    double v = 0.0;
    if (k < K8) {
#pragma unroll
      for (int wI = 0; wI < VL_THREADS / 32; ++wI) v += red_cs[wI * K8M + k];
      zp[k] = v;
    } else {
#pragma unroll
      for (int wI = 0; wI < VL_THREADS / 32; ++wI) v += red_sc[wI][k - K8];
      zp[K8rec + (k - K8)] = v;
    }
  }

  GK_TICK(12);
  return true;
}

// one read tile of the height the window is laid out for
template <int KTM, int MODE, bool FINE_OK = true>
__device__ __forceinline__ bool vl_cluster_tile(const VlProblem* __restrict__ probs, const int2 tm,
                                                const int stage_doubles, VlState& sS) {
  if constexpr (FINE_OK) {
    if (probs[tm.x].zt_reads == 32) return vl_cluster_tile_mt<KTM, MODE, 1>(probs, tm, stage_doubles, sS);
  }
  return vl_cluster_tile_mt<KTM, MODE, 2>(probs, tm, stage_doubles, sS);
}

// The read tile of a restart that finishes last closes the update: fixed-order sums of the per-tile
// records, then the scalar update (no separate launch, no host decision).  R = the state before the update.
template <int MODE>
__device__ __forceinline__ void vl_cluster_close(const VlProblem& P, const VlState& R, int* __restrict__ done_count,
                                                 int4* __restrict__ pstat_entry) {
  static_assert(VL_THREADS == VL_MAX_CLUSTERS, "one thread per cluster");
  __shared__ VlScalarSmem sm;
  const int tid = threadIdx.x;
  const int K8 = 8 * R.ktl, K8rec = 8 * P.KT, ZS = K8rec + VL_ZP_EXTRA;
  for (int c = tid; c < K8 + 7; c += VL_THREADS) {
    if (c < K8) sm.cs[c] = vl_ordered_sum(P.zpart + c, ZS, P.n_ztiles);
    else if (c < K8 + 4) sm.sc[c - K8] = vl_ordered_sum(P.zpart + K8rec + (c - K8), ZS, P.n_ztiles);
    else sm.sc[4 + (c - K8 - 4)] = vl_ordered_sum(P.hpart + (c - K8 - 4), VL_HP_STRIDE, P.n_htiles);
  }
  vl_scalar_update<0>(P, R, P.st, sm, tid, done_count, pstat_entry);
}

// one tile of a per-phase launch (vl_cluster_kernel): the last tile of a restart is found with a ticket
template <int KTM, int MODE>
__device__ __forceinline__ void vl_cluster_body(const VlProblem* __restrict__ probs, const int2 tm,
                                                int* __restrict__ done_count, int4* __restrict__ pstat) {
  __shared__ VlState sS;
  __shared__ int s_last;
  if (!vl_cluster_tile<KTM, MODE>(probs, tm, 0, sS)) return;
  const VlProblem& P = probs[tm.x];
  VlState* Sw = P.st;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&Sw->ticket, 1) == P.n_ztiles - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) Sw->ticket = 0;
  vl_cluster_close<MODE>(P, sS, done_count, pstat + tm.x);
}

template <int KTM, int MODE>
__global__ void __launch_bounds__(VL_THREADS, vl_min_ctas(KTM))
vl_cluster_kernel(const VlProblem* __restrict__ probs, const int2* __restrict__ tiles,
                  int* __restrict__ done_count, int4* __restrict__ pstat) {
  vl_cluster_body<KTM, MODE>(probs, tiles[blockIdx.x], done_count, pstat);
}

// =========================================================================================
// A whole chunk of updates of one kernel class in ONE launch.  Every scheduled restart has a
// scheduling word (VlCtl) that names its current phase -- the haplotype tiles of update k of the
// chunk, then its read tiles -- and the next tile of it to hand out.  A block is a worker for ONE
// tile: it claims whichever tile is open with one atomicAdd on that word (polling while none is),
// runs it, counts it done and exits; the block that finishes the last tile of a phase opens the
// next one.  The launch has as many blocks as the chunk can have tiles; blocks that find every
// restart gone exit at once.  No block ever waits for a tile that nobody runs: a phase is opened
// only when everything it reads has been written, with one exception, the haplotype tiles of
// update k + 1, which are opened when the last read tile of update k has arrived and wait (after
// their operand loop) for the scalar update that this very tile's CTA is running.  Restarts
// advance independently of each other; a restart that leaves its loop, finishes the chunk, or
// whose layout outgrows the kernel class is marked VL_PHASE_LEFT.  Nothing depends on the order in
// which the hardware dispatches blocks.  (One tile per block rather than a loop of tiles per
// block: inside a loop the compiler keeps the constants of both tile kinds in registers across
// iterations and spills in the DMMA loops.)
// =========================================================================================
// Warp 0 of a block: find a restart with a tile to hand out and claim it.  task = (problem, phase,
// tile index); problem < 0 when every restart of the launch has left the chunk.
__device__ __forceinline__ void vl_sched_claim(const int4* __restrict__ chunk_base, const int* __restrict__ rlist,
                                               const int n_r, VlCtl* ctl, const int* n_live, int* task) {
  const int lane = threadIdx.x & 31;
  unsigned idle = 0;
  const int cursor = (int)(blockIdx.x % (unsigned)n_r);     // blocks start their search at different restarts
  for (;;) {
    int got_p = -1;
    unsigned got_phase = 0, got_idx = 0;
    for (int base = 0; base < n_r && got_p < 0; base += 32) {
      const int slot = base + lane;
      int p = -1;
      bool open = false;
      int4 cnt = make_int4(0, 0, 0, 0);          // the restart's tile counts travel with the scheduling data: one latency
      if (slot < n_r) {
        int at = cursor + slot;
        if (at >= n_r) at -= n_r;
        p = rlist[at];
        cnt = __ldg(chunk_base + p);
        const unsigned long long w = *reinterpret_cast<const volatile unsigned long long*>(&ctl[p].claim);
        const unsigned phase = (unsigned)(w >> 32), idx = (unsigned)w;
        open = phase != VL_PHASE_LEFT && idx < (unsigned)((phase & 1u) ? cnt.z : cnt.y);
      }
      unsigned cand = __ballot_sync(0xffffffffu, open);
      while (cand && got_p < 0) {
        const int src = __ffs(cand) - 1;
        cand &= cand - 1;
        unsigned long long old = 0;
        if (lane == src) old = atomicAdd(&ctl[p].claim, 1ull);
        old = __shfl_sync(0xffffffffu, old, src);
        const int pp = __shfl_sync(0xffffffffu, p, src);
        const int nh = __shfl_sync(0xffffffffu, cnt.y, src), nz = __shfl_sync(0xffffffffu, cnt.z, src);
        const unsigned phase = (unsigned)(old >> 32), idx = (unsigned)old;
        if (phase != VL_PHASE_LEFT && idx < (unsigned)((phase & 1u) ? nz : nh)) {
          got_p = pp; got_phase = phase; got_idx = idx;
        }
      }
    }
    if (got_p >= 0) {
      if (lane == 0) { task[0] = got_p; task[1] = (int)got_phase; task[2] = (int)got_idx; }
      return;
    }
    if (*reinterpret_cast<const volatile int*>(n_live) <= 0) {
      if (lane == 0) task[0] = -1;
      return;
    }
    if (++idle > (1u << 26)) __trap();           // a long time without any tile to run although restarts are live: an error, not a hang
    __nanosleep(idle < 4096 ? 40 : 1000);
  }
}

template <int KTM, int MODE, bool SPLIT>
__global__ void __launch_bounds__(VL_THREADS, vl_min_ctas(KTM))
vl_sched_kernel(const VlProblem* __restrict__ probs, const int* __restrict__ rlist, const int n_r, const int n_iter,
                const int4* __restrict__ chunk_base, VlCtl* ctl, int* n_live, int* __restrict__ done_count,
                int4* __restrict__ pstat, const int stage_doubles) {
  __shared__ int s_task[4];
  __shared__ int s_flag;
  if (threadIdx.x < 32) vl_sched_claim(chunk_base, rlist, n_r, ctl, n_live, s_task);
  __syncthreads();
  if (s_task[0] < 0) return;
  GK_DECL(s_task[0] == 0 && s_task[2] == 0);
  // (the task is read again from shared memory after the tile, so that nothing of the scheduler is live in it)
  if ((s_task[1] & 1) == 0) {
    // ---- haplotype tile of update k ----
    {
      const int p = s_task[0], idx = s_task[2], k = s_task[1] >> 1;
      const VlProblem& P = probs[p];
      const int t = idx % P.n_htiles, sp = idx / P.n_htiles;
      const int2 tm = make_int2(p, P.n_hsplit > 1 ? (t | (sp << 20) | (P.n_hsplit << 24)) : t);
      vl_haplo_body<KTM, MODE, SPLIT>(probs, tm, stage_doubles, true, chunk_base[p].x + k);
    }
#ifdef VL_RS_PROFILE
    gk_t_ = clock64();
#endif
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int p = s_task[0];
      const unsigned phase = (unsigned)s_task[1];
      const int k = (int)(phase >> 1);
      const VlProblem& P = probs[p];
      VlCtl* C = ctl + p;
      // C->done counts the tiles of the whole chunk: this phase is complete when it reaches (k + 1) nh + k nz
      if (atomicAdd(&C->done, 1) + 1 == (k + 1) * P.n_htiles * P.n_hsplit + k * P.n_ztiles) {
        // last tile of the phase: the tiles that ran have seen the published state of this update, so it is final here
        __threadfence();
        const bool live = *reinterpret_cast<const volatile int*>(&P.st->active) != 0 &&
                          *reinterpret_cast<const volatile int*>(&P.st->ktl) <= KTM;
        if (!live) {                       // left its loop, or the layout grew past this kernel class: the host reschedules it
          atomicExch(&C->claim, (unsigned long long)VL_PHASE_LEFT << 32);
          __threadfence();
          atomicSub(n_live, 1);
        } else {
          atomicExch(&C->claim, (unsigned long long)(phase + 1u) << 32);
        }
      }
    }
    GK_TICK(7);
    return;
  }
  // ---- read tile of update k ----
  __shared__ VlState sZ;                     // the state before the update
  asm volatile("fence.proxy.async;\n" ::: "memory");     // its bulk copies read what other CTAs wrote
  const bool ran = vl_cluster_tile<KTM, MODE, SPLIT>(probs, make_int2(s_task[0], s_task[2]), stage_doubles, sZ);
  __threadfence();
  __syncthreads();
  const int p = s_task[0];
  const unsigned phase = (unsigned)s_task[1];
  const int k = (int)(phase >> 1);
  const VlProblem& P = probs[p];
  VlCtl* C = ctl + p;
  if (threadIdx.x == 0) s_flag = (atomicAdd(&C->done, 1) + 1 == (k + 1) * (P.n_htiles * P.n_hsplit + P.n_ztiles)) ? 1 : 0;
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  const bool more = k + 1 < n_iter;
  // the haplotype tiles of the next update may start their operand loops while this CTA does the
  // scalar update (they read w z, which is complete, and wait for the published state afterwards)
  if (threadIdx.x == 0 && ran && more) atomicExch(&C->claim, (unsigned long long)(phase + 1u) << 32);
#ifdef VL_RS_PROFILE
  const bool gk_last_ = p == 0 && threadIdx.x == 0;
  long long gk_t2_ = clock64();
#endif
  if (ran) vl_cluster_close<MODE>(P, sZ, done_count, pstat + p);
#ifdef VL_RS_PROFILE
  if (gk_last_) { atomicAdd(&vl_gk_cycles[13], (unsigned long long)(clock64() - gk_t2_)); atomicAdd(&vl_gk_cycles[17], 1ull); }
#endif
  if (!ran || !more) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicExch(&C->claim, (unsigned long long)VL_PHASE_LEFT << 32);
      __threadfence();
      atomicSub(n_live, 1);
    }
  }
}

// The same scheduler with persistent workers: a block runs tile after tile until every restart of
// the launch has left the chunk, so a launch occupies only as many SM slots as it is given and
// launches of different kernel classes share the GPU side by side.  The tiles are functions of their
// own here: inlined into the loop, the compiler allocates the registers of both tile kinds and of the
// loop together and spills in the DMMA loops.
template <int KTM, int MODE, bool SPLIT>
__device__ __noinline__ void vl_haplo_task(const VlProblem* __restrict__ probs, const int2 tm, const int stage_doubles,
                                           const int pub_want) {
  vl_haplo_body<KTM, MODE, SPLIT>(probs, tm, stage_doubles, true, pub_want);
}
template <int KTM, int MODE, bool SPLIT>
__device__ __noinline__ bool vl_cluster_task(const VlProblem* __restrict__ probs, const int2 tm, const int stage_doubles,
                                             VlState* sS) {
  return vl_cluster_tile<KTM, MODE, SPLIT>(probs, tm, stage_doubles, *sS);
}
template <int MODE>
__device__ __noinline__ void vl_close_task(const VlProblem* P, const VlState* R, int* done_count, int4* pstat_entry) {
  vl_cluster_close<MODE>(*P, *R, done_count, pstat_entry);
}

template <int KTM, int MODE, bool SPLIT>
__global__ void __launch_bounds__(VL_THREADS, vl_min_ctas(KTM))
vl_sched_persistent_kernel(const VlProblem* __restrict__ probs, const int* __restrict__ rlist, const int n_r, const int n_iter,
                           const int4* __restrict__ chunk_base, VlCtl* ctl, int* n_live, int* __restrict__ done_count,
                           int4* __restrict__ pstat, const int stage_doubles) {
  __shared__ int s_task[4];
  __shared__ int s_flag;
  __shared__ VlState sZ;
  for (;;) {
    if (threadIdx.x < 32) vl_sched_claim(chunk_base, rlist, n_r, ctl, n_live, s_task);
    __syncthreads();
    if (s_task[0] < 0) return;
    if ((s_task[1] & 1) == 0) {
      {
        const int p = s_task[0], idx = s_task[2], k = s_task[1] >> 1;
        const VlProblem& P = probs[p];
        const int t = idx % P.n_htiles, sp = idx / P.n_htiles;
        vl_haplo_task<KTM, MODE, SPLIT>(probs, make_int2(p, P.n_hsplit > 1 ? (t | (sp << 20) | (P.n_hsplit << 24)) : t),
                                        stage_doubles, chunk_base[p].x + k);
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) {
        const int p = s_task[0];
        const unsigned phase = (unsigned)s_task[1];
        const int k = (int)(phase >> 1);
        const VlProblem& P = probs[p];
        VlCtl* C = ctl + p;
        const int nh = P.n_htiles * P.n_hsplit, nz = P.n_ztiles;
        if (atomicAdd(&C->done, 1) + 1 == (k + 1) * nh + k * nz) {
          __threadfence();
          const bool live = *reinterpret_cast<const volatile int*>(&P.st->active) != 0 &&
                            *reinterpret_cast<const volatile int*>(&P.st->ktl) <= KTM;
          if (!live) {
            atomicExch(&C->claim, (unsigned long long)VL_PHASE_LEFT << 32);
            __threadfence();
            atomicSub(n_live, 1);
          } else {
            atomicExch(&C->claim, (unsigned long long)(phase + 1u) << 32);
          }
        }
      }
      __syncthreads();                       // every thread has read the task before the next claim rewrites it
    } else {
      asm volatile("fence.proxy.async;\n" ::: "memory");
      const bool ran = vl_cluster_task<KTM, MODE, SPLIT>(probs, make_int2(s_task[0], s_task[2]), stage_doubles, &sZ);
      __threadfence();
      __syncthreads();
      const int p = s_task[0];
      const unsigned phase = (unsigned)s_task[1];
      const int k = (int)(phase >> 1);
      const VlProblem& P = probs[p];
      VlCtl* C = ctl + p;
      if (threadIdx.x == 0) s_flag = (atomicAdd(&C->done, 1) + 1 == (k + 1) * (P.n_htiles * P.n_hsplit + P.n_ztiles)) ? 1 : 0;
      __syncthreads();
      if (s_flag) {
        __threadfence();
        const bool more = k + 1 < n_iter;
        if (threadIdx.x == 0 && ran && more) atomicExch(&C->claim, (unsigned long long)(phase + 1u) << 32);
        if (ran) vl_close_task<MODE>(&P, &sZ, done_count, pstat + p);
        if (!ran || !more) {
          __syncthreads();
          if (threadIdx.x == 0) {
            __threadfence();
            atomicExch(&C->claim, (unsigned long long)VL_PHASE_LEFT << 32);
            __threadfence();
            atomicSub(n_live, 1);
          }
        }
      }
      __syncthreads();
    }
  }
}

// One thread per problem: state_init_dict of draw_init_state (initialization.py:6-43), the
// first refresh of the digamma sums (cavi.py:122-128 at iter = 0) and the identity layout.
__global__ void vl_init_kernel(const VlProblem* __restrict__ probs, int n_probs, int4* __restrict__ pstat) {
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
  for (int k = 0; k < K8; ++k) {
    P.alpha[k] = (k < K) ? alpha0 : 0.0;
    P.mlp[k] = (k < K) ? lp0 : 0.0;
    P.mlp_lay[k] = (k < K) ? lp0 : 0.0;
    P.lay[k] = (k < K) ? k : -1;
    P.lay_z[k] = (k < K) ? k : -1;
    P.gat[k] = (k < K) ? k : 0;
  }
  S->lnB0 = K * vl_lgamma(alpha0) - vl_lgamma(sum_alpha);
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
  S->ktl = P.KT; S->ktl_z = P.KT; S->nlay = K; S->ghost = -1; S->ghost_mult = 0.0; S->stalled = 0;
  S->nlay_z = K; S->ghost_z = -1; S->ticket = 0; S->pub = 0;
  S->ktring[0] = P.KT; S->ktring[1] = P.KT;
  pstat[p] = make_int4(1, P.KT, K, 0);
}

// Dense operand of one window from the raw arrays: Dm[n,j] = d where read n carries base b_j at
// position l_j (d = log theta - log((1-theta)/(B-1)) for uqs, 1 for lep), 0 elsewhere (other
// base, 'N', padding).  One thread per (read, column) of a tile.
__global__ void vl_prepare_dm_kernel(const uint8_t* __restrict__ codes, const double* __restrict__ lt,
                                     const double* __restrict__ lo, const int* __restrict__ col_lb,
                                     double* __restrict__ dm, int N, int L, int Npad, int mode) {
  const int ct = blockIdx.y;
  const int n = blockIdx.x * (blockDim.x / VL_PT) + threadIdx.x / VL_PT;
  const int p = threadIdx.x % VL_PT;
  if (n >= Npad) return;
  const int lb = col_lb[ct * VL_PT + p];
  double v = 0.0;
  if (n < N && lb >= 0) {
    const int l = lb / VL_B, b = lb - l * VL_B;
    const size_t off = (size_t)n * L + l;
    if (codes[off] == (unsigned)b) v = (mode == VL_MODE_UQS) ? (lt[off] - lo[off]) : 1.0;
  }
  dm[((size_t)ct * Npad + n) * VL_PT + vl_dm_col(n, p)] = v;
}

// per read: number of called positions (n_non_N of lep preparation.py:13) and, for uqs, the
// sum of log theta over them
__global__ void vl_prepare_rows_kernel(const uint8_t* __restrict__ codes, const double* __restrict__ lt,
                                       double* __restrict__ mcount, double* __restrict__ ltsum,
                                       int N, int L, int Npad, int mode) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= Npad) return;
  const int lane = threadIdx.x & 31;
  double cnt = 0.0, s = 0.0;
  if (n < N)
    for (int l = lane; l < L; l += 32) {
      const size_t off = (size_t)n * L + l;
      if (codes[off] < VL_B) {
        cnt += 1.0;
        if (mode == VL_MODE_UQS) s += lt[off];
      }
    }
  cnt = warp_sum(cnt); s = warp_sum(s);
  if (lane == 0) { mcount[n] = cnt; ltsum[n] = s; }
}

__device__ __forceinline__ void build_slot_table(const VlProblem& P, int* slot_of) {
  // cluster -> slot of the layout z and h are stored in; dead clusters share the ghost slot
  const VlState* S = P.st;
  for (int k = threadIdx.x; k < P.K; k += blockDim.x) slot_of[k] = S->ghost_z;
  __syncthreads();
  for (int s = threadIdx.x; s < S->nlay_z; s += blockDim.x)
    if (s != S->ghost_z) slot_of[P.lay_z[s]] = s;
  __syncthreads();
}

// mean_haplo [Ls][B][KS] in layout order -> reference layout (K, L, B)
__global__ void vl_export_haplo_kernel(const VlProblem* __restrict__ prob, double* __restrict__ out) {
  __shared__ int slot_of[VL_MAX_CLUSTERS];
  const VlProblem& P = *prob;
  build_slot_table(P, slot_of);
  const int KS = vl_ks(P.st->ktl_z);
  const size_t total = (size_t)P.K * P.L * VL_B;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i % VL_B);
    const int l = (int)((i / VL_B) % P.L);
    const int k = (int)(i / ((size_t)VL_B * P.L));
    out[i] = P.h[((size_t)l * VL_B + b) * KS + slot_of[k]];
  }
}

// mean_cluster [Npad][KS] in layout order -> (N, K)
__global__ void vl_export_cluster_kernel(const VlProblem* __restrict__ prob, double* __restrict__ out) {
  __shared__ int slot_of[VL_MAX_CLUSTERS];
  const VlProblem& P = *prob;
  build_slot_table(P, slot_of);
  const int KS = vl_ks(P.st->ktl_z);
  const size_t total = (size_t)P.N * P.K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % P.K);
    const size_t n = i / P.K;
    out[i] = P.z[n * KS + slot_of[k]];
  }
}

// merge_cluster_assignments (analyze_results.py:160-167): members added in ascending k;
// output rows have the stride of an identity layout over G groups
__global__ void vl_merge_kernel(const VlProblem* __restrict__ prob, double* __restrict__ zm,
                                double* __restrict__ wzm, const int* __restrict__ group_of, int G, int KSg) {
  __shared__ int slot_of[VL_MAX_CLUSTERS];
  const VlProblem& P = *prob;
  build_slot_table(P, slot_of);
  const int KS = vl_ks(P.st->ktl_z);
  const size_t total = (size_t)P.Npad * KSg;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t n = i / KSg;
    const int j = (int)(i % KSg);
    double s = 0.0;
    if (j < G && n < (size_t)P.N)
      for (int k = 0; k < P.K; ++k) if (group_of[k] == j) s += P.z[n * KS + slot_of[k]];
    zm[i] = s;
    wzm[i] = P.w[n] * s;
  }
}

// w z of a responsibility matrix that arrived from the host (initial state, teacher forcing)
__global__ void vl_scale_rows_kernel(const double* __restrict__ z, const double* __restrict__ w,
                                     double* __restrict__ wz, int Npad, int KS) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)Npad * KS) return;
  wz[i] = w[i / KS] * z[i];
}

// initial responsibilities as the host holds them ([N][K], contiguous) -> z [Npad][KS] and w z
__global__ void vl_init_rows_kernel(const double* __restrict__ raw, const double* __restrict__ w, double* __restrict__ z,
                                    double* __restrict__ wz, int N, int K, int Npad, int KS) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)Npad * KS) return;
  const int n = (int)(i / KS), k = (int)(i - (size_t)n * KS);
  const double v = (n < N && k < K) ? raw[(size_t)n * K + k] : 0.0;
  z[i] = v;
  wz[i] = w[n] * v;
}

// test hook: the softmax exponential on arbitrary non-positive inputs
__global__ void vl_exp_test_kernel(const double* __restrict__ x, double* __restrict__ out, int n) {
  const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n) return;
  const double xs[2] = {x[i], (i + 1 < n) ? x[i + 1] : 0.0};
  double ev[2];
  vl_exp_nonpos<2>(xs, ev);
  out[i] = ev[0];
  if (i + 1 < n) out[i + 1] = ev[1];
}

// roofline denominator: the FP64 tensor instruction of the contraction kernels (DMMA m8n8k4) in a
// register-resident loop with eight independent accumulators per warp
__global__ void __launch_bounds__(256) vl_dmma_peak_kernel(double* __restrict__ out, int iters, double a0, double b0) {
  double acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  const double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) vl_dmma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KTM>
constexpr size_t smem_haplo() {
  return (size_t)vl_haplo_doubles(KTM) * sizeof(double);
}
template <int KTM>
constexpr size_t smem_cluster() {
  return ((size_t)vl_zstages(KTM) * (VL_ZT_READS * VL_PT + VL_PT * vl_ks(KTM)) + (VL_THREADS / 32) * 8 * KTM) * sizeof(double);
}

template <int KTM, int MODE, bool SPLIT>
cudaError_t launch_haplo(const VlProblem* probs, const int2* tiles, int n, cudaStream_t stream) {
  // the post-loop tables T [16][K8] + Cm [16][5][K8] reuse the stage buffers
  static_assert(smem_haplo<KTM>() >= (size_t)VL_PT * 6 * 8 * KTM * sizeof(double), "haplo smem union");
  cudaError_t e = cudaFuncSetAttribute(vl_haplo_kernel<KTM, MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem_haplo<KTM>());
  if (e != cudaSuccess) return e;
  vl_haplo_kernel<KTM, MODE, SPLIT><<<n, VL_THREADS, smem_haplo<KTM>(), stream>>>(probs, tiles);
  return cudaGetLastError();
}
template <int KTM, int MODE, bool SPLIT>
cudaError_t launch_cluster(const VlProblem* probs, const int2* tiles, int n, int* done_count, int4* pstat,
                           cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(vl_cluster_kernel<KTM, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem_cluster<KTM>());
  if (e != cudaSuccess) return e;
  vl_cluster_kernel<KTM, MODE><<<n, VL_THREADS, smem_cluster<KTM>(), stream>>>(probs, tiles, done_count, pstat);
  return cudaGetLastError();
}

// Shared memory of a worker: sized for residency, or -- when the launch has few tiles in flight, the SMs
// are mostly idle and one restart's update is a chain of latencies -- enough to have (nearly) all
// operand stages of a tile in flight at once.  Same arithmetic either way.
template <int KTM>
size_t sched_smem(int deep, int* stage_doubles) {
  size_t smem = std::max(smem_haplo<KTM>(), smem_cluster<KTM>());
  *stage_doubles = 0;
  // deep = 1: two CTAs per SM; deep = 2: the launch has at most one tile per SM in flight, every CTA gets an SM's shared memory
  static const size_t dev_kb = std::getenv("VL_DEEP_KB") ? (size_t)std::atoi(std::getenv("VL_DEEP_KB")) : 0;   // development aid
  const size_t big = (deep ? (dev_kb ? dev_kb : (deep >= 2 ? (size_t)200 : (size_t)104)) : (size_t)0) * 1024;
  if (big > smem) {
    *stage_doubles = (int)((big - (VL_THREADS / 32) * 8 * KTM * sizeof(double)) / sizeof(double));
    smem = big;
  }
  return smem;
}

template <int KTM, int MODE, bool SPLIT>
cudaError_t sched_occupancy(int deep, int persistent, int* ctas_per_sm, int* smem_bytes, int* regs) {
  int sd = 0;
  const size_t smem = sched_smem<KTM>(deep, &sd);
  const void* fn = persistent ? (const void*)vl_sched_persistent_kernel<KTM, MODE, SPLIT> : (const void*)vl_sched_kernel<KTM, MODE, SPLIT>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(VL_ATTR_KB * 1024));
  if (e != cudaSuccess) return e;
  cudaFuncAttributes at;
  e = cudaFuncGetAttributes(&at, fn);
  if (e != cudaSuccess) return e;
  *regs = at.numRegs;
  *smem_bytes = (int)(smem + at.sharedSizeBytes);
  if (persistent) return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, vl_sched_persistent_kernel<KTM, MODE, SPLIT>, VL_THREADS, smem);
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, vl_sched_kernel<KTM, MODE, SPLIT>, VL_THREADS, smem);
}

template <int KTM, int MODE, bool SPLIT>
cudaError_t launch_sched(const VlProblem* probs, const int* rlist, int n_r, int n_updates, const int4* chunk_base,
                         VlCtl* ctl, int* n_live, int* done_count, int4* pstat, long long n_blocks, int deep, int persistent,
                         cudaStream_t stream) {
  int stage_doubles = 0;
  const size_t smem = sched_smem<KTM>(deep, &stage_doubles);
  if (n_blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
  if (persistent) {
    cudaError_t e = cudaFuncSetAttribute(vl_sched_persistent_kernel<KTM, MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(VL_ATTR_KB * 1024));
    if (e != cudaSuccess) return e;
    vl_sched_persistent_kernel<KTM, MODE, SPLIT><<<(unsigned)n_blocks, VL_THREADS, smem, stream>>>(probs, rlist, n_r, n_updates, chunk_base,
                                                                                            ctl, n_live, done_count, pstat, stage_doubles);
    return cudaGetLastError();
  }
  cudaError_t e = cudaFuncSetAttribute(vl_sched_kernel<KTM, MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(VL_ATTR_KB * 1024));
  if (e != cudaSuccess) return e;
  vl_sched_kernel<KTM, MODE, SPLIT><<<(unsigned)n_blocks, VL_THREADS, smem, stream>>>(probs, rlist, n_r, n_updates, chunk_base, ctl,
                                                                                 n_live, done_count, pstat, stage_doubles);
  return cudaGetLastError();
}

}  // namespace

int vl_class_of_tiles(int kt) { return kt <= 2 ? 0 : (kt <= 4 ? 1 : (kt <= 8 ? 2 : 3)); }
int vl_class_capacity(int cls) { return 2 << cls; }

#define VL_DISPATCH(FN, ...)                                                             \
  switch ((cls * 2 + (mode == VL_MODE_LEP ? 1 : 0)) * 2 + (split ? 1 : 0)) {               \
    case 0: return FN<2, VL_MODE_UQS, false>(__VA_ARGS__);                               \
    case 1: return FN<2, VL_MODE_UQS, true>(__VA_ARGS__);                                \
    case 2: return FN<2, VL_MODE_LEP, false>(__VA_ARGS__);                               \
    case 3: return FN<2, VL_MODE_LEP, true>(__VA_ARGS__);                                \
    case 4: return FN<4, VL_MODE_UQS, false>(__VA_ARGS__);                               \
    case 5: return FN<4, VL_MODE_UQS, true>(__VA_ARGS__);                                \
    case 6: return FN<4, VL_MODE_LEP, false>(__VA_ARGS__);                               \
    case 7: return FN<4, VL_MODE_LEP, true>(__VA_ARGS__);                                \
    case 8: return FN<8, VL_MODE_UQS, false>(__VA_ARGS__);                               \
    case 9: return FN<8, VL_MODE_UQS, true>(__VA_ARGS__);                                \
    case 10: return FN<8, VL_MODE_LEP, false>(__VA_ARGS__);                              \
    case 11: return FN<8, VL_MODE_LEP, true>(__VA_ARGS__);                               \
    case 12: return FN<16, VL_MODE_UQS, false>(__VA_ARGS__);                             \
    case 13: return FN<16, VL_MODE_UQS, true>(__VA_ARGS__);                              \
    case 14: return FN<16, VL_MODE_LEP, false>(__VA_ARGS__);                             \
    case 15: return FN<16, VL_MODE_LEP, true>(__VA_ARGS__);                              \
    default: return cudaErrorInvalidValue;                                               \
  }

cudaError_t vl_sched_occupancy(int cls, int mode, int split, int deep, int persistent, int* ctas_per_sm, int* smem_bytes, int* regs) {
  VL_DISPATCH(sched_occupancy, deep, persistent, ctas_per_sm, smem_bytes, regs)
}

cudaError_t vl_launch_haplo(int cls, int mode, int split, const VlProblem* probs, const int2* tiles, int n_tiles,
                            cudaStream_t stream) {
  if (n_tiles <= 0) return cudaSuccess;
  VL_DISPATCH(launch_haplo, probs, tiles, n_tiles, stream)
}

cudaError_t vl_launch_cluster(int cls, int mode, const VlProblem* probs, const int2* tiles, int n_tiles,
                              int* done_count, int4* pstat, cudaStream_t stream) {
  const int split = 0;
  if (n_tiles <= 0) return cudaSuccess;
  VL_DISPATCH(launch_cluster, probs, tiles, n_tiles, done_count, pstat, stream)
}

cudaError_t vl_launch_sched(int cls, int mode, int split, const VlProblem* probs, const int* rlist, int n_r, int n_updates,
                            const int4* chunk_base, VlCtl* ctl, int* n_live, int* done_count, int4* pstat, long long n_blocks,
                            int deep, int persistent, cudaStream_t stream) {
  if (n_r <= 0 || n_updates <= 0 || n_blocks <= 0) return cudaSuccess;
  VL_DISPATCH(launch_sched, probs, rlist, n_r, n_updates, chunk_base, ctl, n_live, done_count, pstat, n_blocks, deep, persistent, stream)
}

cudaError_t vl_launch_init(const VlProblem* probs, int n_probs, int4* pstat, cudaStream_t stream) {
  if (n_probs > 0) vl_init_kernel<<<(n_probs + 127) / 128, 128, 0, stream>>>(probs, n_probs, pstat);
  return cudaGetLastError();
}

cudaError_t vl_launch_prepare(const uint8_t* codes, const double* lt, const double* lo, const int* col_lb,
                              double* dm, double* mcount, double* ltsum, int N, int L, int Npad, int NCs,
                              int mode, cudaStream_t stream) {
  const dim3 grid((Npad + 15) / 16, NCs / VL_PT);
  vl_prepare_dm_kernel<<<grid, 256, 0, stream>>>(codes, lt, lo, col_lb, dm, N, L, Npad, mode);
  vl_prepare_rows_kernel<<<(Npad + 3) / 4, 128, 0, stream>>>(codes, lt, mcount, ltsum, N, L, Npad, mode);
  return cudaGetLastError();
}

int vl_debug_general_cycles(unsigned long long* out20, int reset) {
#ifdef VL_RS_PROFILE
  if (cudaMemcpyFromSymbol(out20, vl_gk_cycles, sizeof(vl_gk_cycles)) != cudaSuccess) return 1;
  {   // the scalar-update phases travel in slots 18, 19 (+ stderr dump): development aid only
    unsigned long long su[8];
    if (cudaMemcpyFromSymbol(su, vl_su_cycles, sizeof su) != cudaSuccess) return 1;
    fprintf(stderr, "scalar update phases (cycles, summed): slots %llu | per-cluster %llu | barrier %llu | thread0 %llu | layout %llu | publish %llu\n",
            su[0], su[1], su[2], su[3], su[4], su[5]);
    if (reset) { unsigned long long z8[8] = {}; cudaMemcpyToSymbol(vl_su_cycles, z8, sizeof z8); }
  }
  if (reset) {
    unsigned long long z[20] = {};
    if (cudaMemcpyToSymbol(vl_gk_cycles, z, sizeof z) != cudaSuccess) return 1;
  }
  return 0;
#else
  (void)out20; (void)reset;
  return 2;
#endif
}

cudaError_t vl_launch_scale_rows(const double* z, const double* w, double* wz, int Npad, int KS, cudaStream_t stream) {
  const size_t total = (size_t)Npad * KS;
  vl_scale_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(z, w, wz, Npad, KS);
  return cudaGetLastError();
}

cudaError_t vl_launch_init_rows(const double* raw, const double* w, double* z, double* wz, int N, int K, int Npad, int KS,
                                cudaStream_t stream) {
  const size_t total = (size_t)Npad * KS;
  vl_init_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(raw, w, z, wz, N, K, Npad, KS);
  return cudaGetLastError();
}

cudaError_t vl_launch_dmma_peak(double* out, int n_blocks, int iters, cudaStream_t stream) {
  vl_dmma_peak_kernel<<<n_blocks, 256, 0, stream>>>(out, iters, 1.0, 1e-3);
  return cudaGetLastError();
}

cudaError_t vl_launch_exp_test(const double* x, double* out, int n, cudaStream_t stream) {
  vl_exp_test_kernel<<<(n / 2 + 128) / 128, 128, 0, stream>>>(x, out, n);
  return cudaGetLastError();
}

static unsigned export_grid(size_t total) {
  return (unsigned)std::min<size_t>((total + 255) / 256, 148 * 8);
}

cudaError_t vl_launch_export_haplo(const VlProblem* prob, double* out, size_t total, cudaStream_t stream) {
  vl_export_haplo_kernel<<<export_grid(total), 256, 0, stream>>>(prob, out);
  return cudaGetLastError();
}

cudaError_t vl_launch_export_cluster(const VlProblem* prob, double* out, size_t total, cudaStream_t stream) {
  vl_export_cluster_kernel<<<export_grid(total), 256, 0, stream>>>(prob, out);
  return cudaGetLastError();
}

cudaError_t vl_launch_merge(const VlProblem* prob, double* zm, double* wzm, const int* group_of, int G, int KSg,
                            size_t total, cudaStream_t stream) {
  vl_merge_kernel<<<export_grid(total), 256, 0, stream>>>(prob, zm, wzm, group_of, G, KSg);
  return cudaGetLastError();
}
