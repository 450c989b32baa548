// Resident CAVI kernel for sm_100a: one persistent CTA per (window, restart) problem runs whole
// CAVI iterations on one SM -- no launch and no grid-wide synchronisation between the steps of
// an iteration, and every restart advances at its own pace until it leaves its loop.
//
// Eligible restarts: layout of at most 4 cluster tiles (32 live clusters + ghost), at most 16
// column tiles, operand ring + tables within the 227 KB of shared memory (vl_resident_smem_bytes).
// Everything else, and the first iterations of every restart while most of the K clusters are
// still alive, runs on the general-path kernels (vl_kernels.cu); the two paths share the
// state in global memory and can alternate at chunk boundaries.
//
// One iteration = ONE pass over the dense operand Dm (vl_device.cuh), streamed through a
// shared-memory ring by a dedicated TMA producer warp (cp.async.bulk + full/empty mbarriers):
//   [A] haplotype table of the iteration from T = Dm^T (w z) accumulated during the previous
//       pass (+ sparse entries): softmax over the bases -> h, g (global), gm (shared)
//   [B] per tile of 64 reads:  pass 1  Y = Dm gm  (DMMA, M = read)  -> softmax over the slots
//       -> z (global), w z (shared);  pass 2  T += Dm^T (w z)  (DMMA, M = column) on the SAME
//       operand tile while it is still in the ring
//   [C] fixed-order reduction of the column sums / ELBO pieces, vl_scalar_update (vl_cavi.cuh)
// References: update_eqs.py:5-38 (uqs), lep update_eqs.py:32-51, elbo_eqs.py, cavi.py:120-169.
#include <algorithm>
#include <type_traits>

#include "vl_cavi.cuh"
#include "vl_launch.h"

namespace {

constexpr int RS_WARPS = 8;                       // consumer warps
constexpr int RS_CONSUMERS = 32 * RS_WARPS;
constexpr int RS_THREADS = RS_CONSUMERS + 32;     // + the producer warp
constexpr int RS_KTM = VL_RS_MAX_KT;              // widest layout (8-slot tiles)
constexpr int RS_UNITS = 4;                       // 8-column units of T per warp: 16 column tiles max
constexpr int RS_TILE = VL_ZT_READS * VL_PT;      // doubles per ring stage: 64 reads x 16 columns
constexpr uint32_t RS_TILE_BYTES = RS_TILE * sizeof(double);
constexpr int BAR_ALL = 1, BAR_SCALAR = 2;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(vl_smem_u32(bar)) : "memory");
}

// Per-phase cycle counters of one CTA (blockIdx 0, thread 0), accumulated over a launch; read by
// vl_debug_resident_cycles.  Compiled in only with -DVL_RS_PROFILE.
#ifdef VL_RS_PROFILE
__device__ unsigned long long vl_rs_cycles[16];
#define RS_TICK(slot)                                                     \
  do {                                                                    \
    if (blockIdx.x == 0 && tid == 0) {                                    \
      const long long now_ = clock64();                                   \
      vl_rs_cycles[slot] += (unsigned long long)(now_ - tick_);           \
      tick_ = now_;                                                       \
    }                                                                     \
  } while (0)
#else
#define RS_TICK(slot) do { } while (0)
#endif

struct ResidentShared {
  uint64_t full[VL_RS_MAX_CT], empty[VL_RS_MAX_CT], go;
  VlScalarSmem sc;
  double red_cs[RS_WARPS][8 * RS_KTM];
  double red_sc[RS_WARPS][8];
  int inv[8 * RS_KTM];       // slot of the layout T was accumulated in -> slot of the current one
  VlState st;                // the restart's state as of the start of the current update
  volatile int cmd;          // consumers -> producer: 1 = stream one sweep over the reads, 0 = exit
};

template <int MODE>
// 9 warps: one SM sub-partition hosts three of them, which caps the kernel at 168 registers
__global__ void __launch_bounds__(RS_THREADS, 1)
vl_resident_kernel(const VlProblem* __restrict__ probs, const int* __restrict__ plist, int n_iter,
                   int* __restrict__ done_count, int4* __restrict__ pstat) {
  extern __shared__ __align__(128) unsigned char vl_smem[];
  __shared__ ResidentShared rs;

  const int pid = plist[blockIdx.x];
  const VlProblem& P = probs[pid];
  VlState* S = P.st;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
  const int Npad = P.Npad, NCs = P.NCs, NCT = NCs / VL_PT, NRT = Npad / VL_ZT_READS;
  {
    const int kt0 = max(S->ktl, S->ktl_z);
    if (S->active == 0 || kt0 > RS_KTM || NCT > VL_RS_MAX_CT) return;
  }
  const int KScap = vl_ks(max(S->ktl, S->ktl_z));   // layouts only shrink while the kernel runs
  double* ring = reinterpret_cast<double*>(vl_smem);
  double* gm_s = ring + (size_t)NCT * RS_TILE;       // [NCs][KS]
  double* wz_s = gm_s + (size_t)NCs * KScap;         // [64][KS]

  if (tid == 0) {
    for (int c = 0; c < NCT; ++c) { vl_mbar_init(&rs.full[c], 1); vl_mbar_init(&rs.empty[c], 2); }
    vl_mbar_init(&rs.go, 1);
    vl_fence_mbar_init();
  }
  __syncthreads();

  // ------------------------------------------------------------------ producer warp
  if (warp == RS_WARPS) {
    if (lane == 0) {
      uint32_t use = 0, go_par = 0;
      for (;;) {
        vl_mbar_wait(&rs.go, go_par);
        go_par ^= 1;
        if (rs.cmd == 0) break;
        for (int rt = 0; rt < NRT; ++rt, ++use)
          for (int c = 0; c < NCT; ++c) {
            if (use > 0) vl_mbar_wait(&rs.empty[c], (use - 1) & 1);
            vl_mbar_expect_tx(&rs.full[c], RS_TILE_BYTES);
            vl_tma_load_1d(ring + (size_t)c * RS_TILE, P.dm + ((size_t)c * Npad + (size_t)rt * VL_ZT_READS) * VL_PT,
                           RS_TILE_BYTES, &rs.full[c]);
          }
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumer warps
  const int* __restrict__ gat = P.gat;
  const int* __restrict__ poscol = P.poscol;
  const double* __restrict__ wg = P.w;
  uint32_t use_c = 0;
#ifdef VL_RS_PROFILE
  long long tick_ = clock64();
#endif
  double Tacc[RS_UNITS][RS_KTM][2];
  int ktT = 0;                     // tiles of the layout Tacc was accumulated in

  auto signal_producer = [&](int cmd) {
    if (tid == 0) { rs.cmd = cmd; __threadfence_block(); mbar_arrive(&rs.go); }
  };
  // pass 2 of one read tile: T += Dm^T (w z).  The warp's (up to) four 8-column units advance
  // together: one w z fragment feeds four independent DMMA chains; the ring stages are released
  // at the end.
  auto pass2_t = [&](auto ktc, int KS, bool wait_full) {
    constexpr int KT = decltype(ktc)::value;     // compile-time, see vl_dispatch_tiles
    const double* dtu[RS_UNITS];
    bool vu[RS_UNITS];
#pragma unroll
    for (int u = 0; u < RS_UNITS; ++u) {
      const int mu = warp + RS_WARPS * u;
      vu[u] = mu < 2 * NCT;
      const int c = vu[u] ? (mu >> 1) : 0;
      if (vu[u] && wait_full) vl_mbar_wait(&rs.full[c], use_c & 1);
      dtu[u] = ring + (size_t)c * RS_TILE + (((mu & 1) * 8 + g + 4 * q) & 15) + q * VL_PT;
    }
#pragma unroll 2
    for (int s = 0; s < VL_ZT_READS / 4; ++s) {
      const double* wr = wz_s + (4 * s + q) * KS + g;
      double bf[KT];
#pragma unroll
      for (int i = 0; i < KT; ++i) bf[i] = wr[8 * i];
#pragma unroll
      for (int u = 0; u < RS_UNITS; ++u)
        if (vu[u]) {
          const double a = dtu[u][4 * s * VL_PT];
#pragma unroll
          for (int i = 0; i < KT; ++i) vl_dmma(Tacc[u][i], a, bf[i]);
        }
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int u = 0; u < RS_UNITS; ++u)
        if (vu[u]) mbar_arrive(&rs.empty[(warp + RS_WARPS * u) >> 1]);
    }
  };
  auto pass2 = [&](int kt, int KS, bool wait_full) {
    vl_dispatch_tiles<RS_KTM>(kt, [&](auto ktc) { pass2_t(ktc, KS, wait_full); });
  };
  // pass 1 of one read tile: Y = Dm gm for this warp's 8 reads, two accumulator chains per tile
  auto pass1_t = [&](auto ktc, int KS, int r0, double (&acc)[RS_KTM][2]) {
    constexpr int KT = decltype(ktc)::value;
    double acc_b[KT][2];
#pragma unroll
    for (int i = 0; i < KT; ++i) { acc_b[i][0] = 0.0; acc_b[i][1] = 0.0; }
    for (int c = 0; c < NCT; ++c) {
      vl_mbar_wait(&rs.full[c], use_c & 1);
      const double* dt = ring + (size_t)c * RS_TILE + r0 * VL_PT;
      const double* gt = gm_s + (size_t)c * VL_PT * KS + g;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const double a = dt[(4 * s + q + 4 * (g & 3)) & 15];
        const double* gr = gt + (4 * s + q) * KS;
#pragma unroll
        for (int i = 0; i < KT; ++i) {
          if (s & 1) vl_dmma(acc_b[i], a, gr[8 * i]); else vl_dmma(acc[i], a, gr[8 * i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < KT; ++i) { acc[i][0] += acc_b[i][0]; acc[i][1] += acc_b[i][1]; }
  };

  for (int it = 0; it < n_iter; ++it) {
    vl_bar_sync<BAR_ALL>(RS_CONSUMERS);
    vl_load_state(&rs.st, S, tid);
    vl_bar_sync<BAR_ALL>(RS_CONSUMERS);
    const VlState* S = &rs.st;                // (shadows the global state: reads come from the copy)
    const int kt = S->ktl, ktz = S->ktl_z, nlay = S->nlay, ghost = S->ghost;
    const double gmult = S->ghost_mult;
    const int KS = vl_ks(kt), KSz = vl_ks(ktz), K8 = 8 * kt;

    if (it == 0) {
      // ---- first iteration of this launch: T from the stored z, directly in the current layout
#pragma unroll
      for (int u = 0; u < RS_UNITS; ++u)
#pragma unroll
        for (int i = 0; i < RS_KTM; ++i) { Tacc[u][i][0] = 0.0; Tacc[u][i][1] = 0.0; }
      signal_producer(1);
      for (int rt = 0; rt < NRT; ++rt) {
        for (int idx = tid; idx < VL_ZT_READS * K8; idx += RS_CONSUMERS) {
          const int row = idx / K8, s = idx - row * K8;
          const int n = rt * VL_ZT_READS + row;
          wz_s[row * KS + s] = (s < nlay) ? P.wz[(size_t)n * KSz + gat[s]] : 0.0;
        }
        vl_bar_sync<BAR_ALL>(RS_CONSUMERS);
        pass2(kt, KS, true);
        vl_bar_sync<BAR_ALL>(RS_CONSUMERS);
        ++use_c;
      }
      ktT = kt;
      if (tid < 8 * RS_KTM) rs.inv[tid] = tid;
      RS_TICK(0);
    }

    // ================= [A] haplotype table of this iteration =================
    vl_bar_sync<BAR_ALL>(RS_CONSUMERS);
    double* Ts = ring;                 // [NCs][K8]; the ring is idle between two sweeps
#pragma unroll
    for (int u = 0; u < RS_UNITS; ++u) {
      const int mu = warp + RS_WARPS * u;
      if (mu < 2 * NCT) {
#pragma unroll
        for (int i = 0; i < RS_KTM; ++i)
          if (i < ktT) {
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const int sn = rs.inv[8 * i + 2 * q + jj];
              if (sn >= 0) Ts[(size_t)(8 * mu + g) * K8 + sn] = Tacc[u][i][jj];
            }
          }
      }
#pragma unroll
      for (int i = 0; i < RS_KTM; ++i) { Tacc[u][i][0] = 0.0; Tacc[u][i][1] = 0.0; }
    }
    // padding columns of the operand contribute nothing
    for (int idx = tid; idx < NCs * K8; idx += RS_CONSUMERS) {
      const int j = idx / K8;
      if (P.col_lb[j] < 0) gm_s[(size_t)j * KS + (idx - j * K8)] = 0.0;
    }
    vl_bar_sync<BAR_ALL>(RS_CONSUMERS);
    RS_TICK(1);
    {
      const double mlg0 = S->mlg0, refoff = S->mlg1 - VL_LOG4;
      double cscale = 1.0;
      if (MODE == VL_MODE_LEP) cscale = S->mlt0 - (S->mlt1 - VL_LOG4);
      const int ppw = (K8 <= 16) ? 32 / K8 : 1;          // positions per warp step
      const int sub = lane / K8, slot = lane - sub * K8;
      const int L = P.L;
      double s_ref = 0.0, s_nref = 0.0, s_ent = 0.0;
      // static per-position data (reference base, dense columns, sparse ranges), fetched one
      // step ahead so that its latency overlaps the softmax arithmetic of the current step
      struct PosMeta { int l; unsigned rc; int jc[VL_B]; int ep[VL_B + 1]; };
      auto load_meta = [&](int base) {
        PosMeta m;
        m.l = (sub < ppw && base + sub < L) ? base + sub : -1;
        m.rc = 0;
#pragma unroll
        for (int b = 0; b < VL_B; ++b) { m.jc[b] = -1; m.ep[b] = 0; }
        m.ep[VL_B] = 0;
        if (m.l >= 0) {
          m.rc = P.ref[m.l];
#pragma unroll
          for (int b = 0; b < VL_B; ++b) m.jc[b] = poscol[m.l * VL_B + b];
#pragma unroll
          for (int b = 0; b <= VL_B; ++b) m.ep[b] = P.sp_ptr[m.l * VL_B + b];
        }
        return m;
      };
      const int zcol = (slot < nlay) ? gat[slot] : 0;
      PosMeta nxt = load_meta(warp * ppw);
      for (int base = warp * ppw; base < L; base += RS_WARPS * ppw) {
        const PosMeta cur = nxt;
        nxt = load_meta(base + RS_WARPS * ppw);
        const int l = cur.l;
        if (l < 0) continue;
        const unsigned rc = cur.rc;
        double* hp = P.h + (size_t)l * VL_B * KS + slot;
        double* gp = P.g + (size_t)l * VL_B * KS + slot;
        if (slot < nlay) {
          double lgt[VL_B];
#pragma unroll
          for (int b = 0; b < VL_B; ++b) {
            double cb;
            if (cur.jc[b] >= 0) {
              cb = Ts[(size_t)cur.jc[b] * K8 + slot];
            } else {
              cb = 0.0;
              for (int e = cur.ep[b]; e < cur.ep[b + 1]; ++e)
                cb += P.sp_d[e] * P.wz[(size_t)P.sp_n[e] * KSz + zcol];
            }
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
          const double mu_ = (slot == ghost) ? gmult : 1.0;
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
            hp[b * KS] = hb;
            gp[b * KS] = gb;
            if (cur.jc[b] >= 0) gm_s[(size_t)cur.jc[b] * KS + slot] = gb;
          }
          s_ref += mu_ * r0; s_nref += mu_ * r1; s_ent += mu_ * r2;
        } else {
#pragma unroll
          for (int b = 0; b < VL_B; ++b) {
            hp[b * KS] = 0.0; gp[b * KS] = 0.0;
            if (cur.jc[b] >= 0) gm_s[(size_t)cur.jc[b] * KS + slot] = 0.0;
          }
        }
      }
      s_ref = warp_sum(s_ref); s_nref = warp_sum(s_nref); s_ent = warp_sum(s_ent);
      if (lane == 0) { rs.red_sc[warp][4] = s_ref; rs.red_sc[warp][5] = s_nref; rs.red_sc[warp][6] = s_ent; }
    }
    vl_bar_sync<BAR_ALL>(RS_CONSUMERS);     // Ts consumed, gm_s and the g table complete
    RS_TICK(2);

    // ================= [B] one pass over the reads =================
    signal_producer(1);
    const double* __restrict__ mlpl = P.mlp_lay;
    double b1 = 0.0, b2 = 0.0;
    const double Lf = (double)P.L;
    if (MODE == VL_MODE_LEP) { b1 = S->mlt0; b2 = S->mlt1 - VL_LOG4; }
    double s_data = 0.0, s_ent = 0.0, s_c = 0.0, s_d = 0.0;
    double cs[RS_KTM][2];
#pragma unroll
    for (int i = 0; i < RS_KTM; ++i) { cs[i][0] = 0.0; cs[i][1] = 0.0; }
    const int N = P.N;
    double mlp_r[RS_KTM][2];
#pragma unroll
    for (int i = 0; i < RS_KTM; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int sl = 8 * i + 2 * q + j;
        mlp_r[i][j] = (i < kt && sl < nlay) ? mlpl[sl] : 0.0;
      }
    const int r0 = warp * 8 + g;
    for (int rt = 0; rt < NRT; ++rt) {
      // per-read scalars of this tile: issued now, consumed after the DMMA loop
      const int n = rt * VL_ZT_READS + r0;
      const double wn = wg[n];
      const double mc = P.mcount[n];
      const double ltn = (MODE == VL_MODE_UQS) ? P.ltsum[n] : 0.0;
      const int se0 = P.sr_ptr[n], se1 = P.sr_ptr[n + 1];
      // ---- pass 1 ----
      double acc[RS_KTM][2];
#pragma unroll
      for (int i = 0; i < RS_KTM; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
      vl_dispatch_tiles<RS_KTM>(kt, [&](auto ktc) { pass1_t(ktc, KS, r0, acc); });
      RS_TICK(3);
      for (int e = se0; e < se1; ++e) {
        const double d = P.sr_d[e];
        const double* gp = P.g + (size_t)P.sr_idx[e] * KS + 2 * q;
#pragma unroll
        for (int i = 0; i < RS_KTM; ++i)
          if (i < kt) {
            const double2 gv = *reinterpret_cast<const double2*>(gp + 8 * i);
            acc[i][0] += d * gv.x;
            acc[i][1] += d * gv.y;
          }
      }
      RS_TICK(4);
      // ---- logits, softmax over the slots, reductions, z and w z ----
      {
        const bool valid = n < N;
        const double base = (MODE == VL_MODE_UQS) ? ltn : mc;
        double t[RS_KTM][2];
        double m = neg_inf();
#pragma unroll
        for (int i = 0; i < RS_KTM; ++i)
          if (i < kt) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int s = 8 * i + 2 * q + j;
              const double x = base - acc[i][j];
              acc[i][j] = x;
              double tv = neg_inf();
              if (s < nlay) tv = ((MODE == VL_MODE_UQS) ? x : (b1 * x + b2 * (Lf - x))) + mlp_r[i][j];
              t[i][j] = tv;
              m = fmax(m, tv);
            }
          }
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
        double sum = 0.0, u = 0.0, v = 0.0;
#pragma unroll
        for (int i = 0; i < RS_KTM; ++i)
          if (i < kt) {
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
                v += em * acc[i][j];
                sum += em;
              }
              t[i][j] = e;
            }
          }
        const double psum = sum;
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        const double rsum = 1.0 / sum;
        if (valid) {
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
        double* wrow = wz_s + r0 * KS + 2 * q;
#pragma unroll
        for (int i = 0; i < RS_KTM; ++i)
          if (i < kt) {
            const double z0 = valid ? t[i][0] * rsum : 0.0;
            const double z1 = valid ? t[i][1] * rsum : 0.0;
            *reinterpret_cast<double2*>(zrow + 8 * i) = make_double2(z0, z1);
            *reinterpret_cast<double2*>(wzrow + 8 * i) = make_double2(wn * z0, wn * z1);
            *reinterpret_cast<double2*>(wrow + 8 * i) = make_double2(wn * z0, wn * z1);
            cs[i][0] += wn * z0; cs[i][1] += wn * z1;
          }
      }
      RS_TICK(5);
      vl_bar_sync<BAR_ALL>(RS_CONSUMERS);
      RS_TICK(6);
      // ---- pass 2 on the same operand tile ----
      pass2(kt, KS, false);
      RS_TICK(7);
      vl_bar_sync<BAR_ALL>(RS_CONSUMERS);
      RS_TICK(8);
      ++use_c;
    }
    ktT = kt;

    // ================= [C] reductions and the scalar update =================
#pragma unroll
    for (int i = 0; i < RS_KTM; ++i)
      if (i < kt) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          double v = cs[i][j];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          if (g == 0) rs.red_cs[warp][8 * i + 2 * q + j] = v;
        }
      }
    s_data = warp_sum(s_data); s_ent = warp_sum(s_ent); s_c = warp_sum(s_c); s_d = warp_sum(s_d);
    if (lane == 0) { rs.red_sc[warp][0] = s_data; rs.red_sc[warp][1] = s_ent; rs.red_sc[warp][2] = s_c; rs.red_sc[warp][3] = s_d; }
    vl_bar_sync<BAR_ALL>(RS_CONSUMERS);
    if (tid < K8 + 7) {
      double v = 0.0;
      if (tid < K8) {
#pragma unroll
        for (int wI = 0; wI < RS_WARPS; ++wI) v += rs.red_cs[wI][tid];
        rs.sc.cs[tid] = v;
      } else {
#pragma unroll
        for (int wI = 0; wI < RS_WARPS; ++wI) v += rs.red_sc[wI][tid - K8];
        rs.sc.sc[tid - K8] = v;
      }
    }
    vl_bar_sync<BAR_ALL>(RS_CONSUMERS);
    RS_TICK(9);
    if (tid < VL_MAX_CLUSTERS) vl_scalar_update<BAR_SCALAR>(P, rs.st, P.st, rs.sc, tid, done_count, pstat + pid);
    if (tid < 8 * RS_KTM) rs.inv[tid] = -1;
    vl_bar_sync<BAR_ALL>(RS_CONSUMERS);
    RS_TICK(10);
#ifdef VL_RS_PROFILE
    if (blockIdx.x == 0 && tid == 0) { vl_rs_cycles[14] += 1; vl_rs_cycles[15] += (unsigned long long)kt; }
#endif
    if (rs.sc.stop || P.st->ktl > RS_KTM) break;
    if (tid < P.st->nlay) rs.inv[gat[tid]] = tid;    // current layout <- layout of z and T
  }
  signal_producer(0);
}

}  // namespace

size_t vl_resident_smem_bytes(int n_col_tiles, int kt) {
  const size_t ks = (size_t)vl_ks(kt);
  return ((size_t)n_col_tiles * RS_TILE + (size_t)n_col_tiles * VL_PT * ks + (size_t)VL_ZT_READS * ks) * sizeof(double);
}

int vl_debug_resident_cycles(unsigned long long* out16, int reset) {
#ifdef VL_RS_PROFILE
  if (cudaMemcpyFromSymbol(out16, vl_rs_cycles, sizeof(vl_rs_cycles)) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[16] = {};
    if (cudaMemcpyToSymbol(vl_rs_cycles, z, sizeof z) != cudaSuccess) return 1;
  }
  return 0;
#else
  (void)out16; (void)reset;
  return 2;
#endif
}

size_t vl_resident_smem_limit() { return (size_t)227 * 1024 - sizeof(ResidentShared) - 1024; }

cudaError_t vl_launch_resident(int mode, const VlProblem* probs, const int* plist, int n, int n_iter, size_t smem,
                               int* done_count, int4* pstat, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e;
  if (mode == VL_MODE_UQS) {
    e = cudaFuncSetAttribute(vl_resident_kernel<VL_MODE_UQS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    vl_resident_kernel<VL_MODE_UQS><<<n, RS_THREADS, smem, stream>>>(probs, plist, n_iter, done_count, pstat);
  } else {
    e = cudaFuncSetAttribute(vl_resident_kernel<VL_MODE_LEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    vl_resident_kernel<VL_MODE_LEP><<<n, RS_THREADS, smem, stream>>>(probs, plist, n_iter, done_count, pstat);
  }
  return cudaGetLastError();
}
