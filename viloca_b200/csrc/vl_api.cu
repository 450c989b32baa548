// C ABI of the CAVI engine (include/viloca_b200.h): workspace layout, window analysis
// (singled-out base per position, minority lists), host<->device transfers, the iteration
// driver with its per-chunk kernel-class scheduling, and result export.  No process-global
// state other than a thread-local error string.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/viloca_b200.h"
#include "vl_launch.h"

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) { g_err = msg; return 1; }

#define VL_CUDA(expr)                                                                           \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) {                                                                    \
      char buf_[512];                                                                           \
      snprintf(buf_, sizeof buf_, "CUDA error %s (%s) at %s:%d", cudaGetErrorName(e_),          \
               cudaGetErrorString(e_), __FILE__, __LINE__);                                     \
      return fail(buf_);                                                                        \
    }                                                                                           \
  } while (0)

constexpr size_t kAlign = 256;
inline size_t align_up(size_t x, size_t a = kAlign) { return (x + a - 1) / a * a; }
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct RestartLayout {
  size_t z, wz, h, g, gm, zpart, hpart, tsplit, hsplit_count, alpha, mlp, mlp_lay, lay, gat, lay_z, hist;
};

struct WindowLayout {
  int mode, N, L, K, R, max_iter;
  int Npad, Ls, KT, KS, K8, n_ztiles, n_htiles, n_hsplit, zt_reads;
  double alpha0, conv_thres;
  size_t nnz_cap;
  int NCs;
  size_t dm, w, ltsum, mcount, ref, col_lb, poscol, tile_pos, sr_ptr, sr_idx, sr_d, sp_ptr, sp_n, sp_wd;
  size_t meta_end;      // ref .. meta_end is one contiguous block of small per-window arrays
  size_t arena_off;     // its place in the pinned host arena
  std::vector<RestartLayout> rs;
  int first_problem;
};

struct Layout {
  std::vector<WindowLayout> wins;
  int n_problems = 0;
  size_t data_begin = 0, data_end = 0;   // zero-filled on upload
  size_t probs = 0, states = 0, done = 0, pstat = 0;
  size_t htiles = 0, ztiles = 0, rlist = 0, cbase = 0, ctl = 0, nlive = 0;
  size_t raw_codes = 0, raw_lt = 0, raw_lo = 0;      // staging of one window's raw arrays
  size_t export_h = 0, export_z = 0;
  size_t tmp_z = 0, tmp_wz = 0, tmp_h = 0, tmp_g = 0, tmp_gm = 0, tmp_hpart = 0, tmp_group = 0, tmp_prob = 0,
         tmp_state = 0, tmp_tiles = 0, tmp_lay = 0, tmp_gat = 0, tmp_mlp = 0;
  size_t total = 0;
  size_t n_htiles_total = 0, n_ztiles_total = 0;
  size_t arena_bytes = 0;
};

// What the kernels need to know about the bases of one window (vl_device.cuh): which
// (position, base) pairs are dense columns of the operand, and the remaining entries listed by
// read (ascending position) and by (position, base) (ascending read).
struct WindowAnalysis {
  int NCs = 0;                          // dense columns incl. padding, a multiple of 16
  std::vector<int> col_lb, poscol, tile_pos;
  std::vector<int> sr_ptr, sr_idx, sp_ptr, sp_n;
  std::vector<double> sr_d, sp_wd;
  size_t nnz = 0;                       // sparse entries
};

// Dense columns: the most frequent called base of every position (first on ties; base 0 if no
// read covers the position) and every other pair carried by at least `thr` reads, where thr
// grows with N: a dense column costs O(N) per contraction, a sparse entry one gathered row.
// The columns of a position are consecutive and never straddle a 16-column tile; tiles are filled
// first-fit in position order.
void plan_columns(const vl_window_desc& d, int Ls, WindowAnalysis& A) {
  const int N = d.n_reads, L = d.length;
  std::vector<int> cnt((size_t)L * VL_N_BASES, 0);
  for (int n = 0; n < N; ++n) {
    const uint8_t* row = d.codes + (size_t)n * L;
    for (int l = 0; l < L; ++l)
      if (row[l] < VL_N_BASES) ++cnt[(size_t)l * VL_N_BASES + row[l]];
  }
  const int thr = std::min(std::max(N / 128, 2), 64);
  A.poscol.assign((size_t)Ls * VL_N_BASES, -1);
  // first fit: a position goes to the first tile that still has room for all its columns, so
  // the single-column positions fill the gaps the multi-column ones leave
  std::vector<std::vector<int>> tile_cols, tile_positions;
  size_t first_open = 0, sparse = 0;
  for (int l = 0; l < L; ++l) {
    const int* c = &cnt[(size_t)l * VL_N_BASES];
    int best = 0;
    for (int b = 1; b < VL_N_BASES; ++b) if (c[b] > c[best]) best = b;
    int bases[VL_N_BASES], nb = 0;
    for (int b = 0; b < VL_N_BASES; ++b) {
      if (b == best || c[b] >= thr) bases[nb++] = b;
      else sparse += c[b];
    }
    while (first_open < tile_cols.size() && tile_cols[first_open].size() == VL_PT) ++first_open;
    size_t t = first_open;
    while (t < tile_cols.size() && tile_cols[t].size() + nb > VL_PT) ++t;
    if (t == tile_cols.size()) { tile_cols.emplace_back(); tile_positions.emplace_back(); }
    for (int i = 0; i < nb; ++i) tile_cols[t].push_back(l * VL_N_BASES + bases[i]);
    tile_positions[t].push_back(l);
  }
  A.col_lb.clear(); A.tile_pos.clear();
  for (size_t t = 0; t < tile_cols.size(); ++t) {
    for (int lb : tile_cols[t]) {
      A.poscol[lb] = (int)A.col_lb.size();
      A.col_lb.push_back(lb);
    }
    A.col_lb.resize((t + 1) * VL_PT, -1);
    A.tile_pos.insert(A.tile_pos.end(), tile_positions[t].begin(), tile_positions[t].end());
    A.tile_pos.resize((t + 1) * VL_PT, -1);
  }
  A.NCs = (int)A.col_lb.size();
  A.nnz = sparse;
}

void analyze_window(const vl_window_desc& d, int Npad, int Ls, WindowAnalysis& A) {
  const int N = d.n_reads, L = d.length;
  const bool uqs = d.mode == VL_MODE_UQS;
  plan_columns(d, Ls, A);
  A.sr_ptr.assign((size_t)Npad + 1, 0);
  A.sp_ptr.assign((size_t)Ls * VL_N_BASES + 1, 0);
  A.sr_idx.clear(); A.sr_d.clear();
  for (int n = 0; n < N; ++n) {
    const uint8_t* row = d.codes + (size_t)n * L;
    for (int l = 0; l < L; ++l) {
      const uint8_t c = row[l];
      if (c < VL_N_BASES && A.poscol[(size_t)l * VL_N_BASES + c] < 0) {
        const size_t off = (size_t)n * L + l;
        A.sr_idx.push_back(l * VL_N_BASES + c);
        A.sr_d.push_back(uqs ? d.log_hit[off] - d.log_off[off] : 1.0);
        ++A.sp_ptr[(size_t)l * VL_N_BASES + c + 1];
      }
    }
    A.sr_ptr[n + 1] = (int)A.sr_idx.size();
  }
  for (int n = N; n < Npad; ++n) A.sr_ptr[n + 1] = A.sr_ptr[N];
  A.nnz = A.sr_idx.size();
  for (size_t i = 1; i < A.sp_ptr.size(); ++i) A.sp_ptr[i] += A.sp_ptr[i - 1];
  A.sp_n.assign(A.nnz, 0);
  A.sp_wd.assign(A.nnz, 0.0);
  std::vector<int> fill(A.sp_ptr.begin(), A.sp_ptr.end() - 1);
  for (int n = 0; n < N; ++n) {
    for (int e = A.sr_ptr[n]; e < A.sr_ptr[n + 1]; ++e) {
      const int at = fill[A.sr_idx[e]]++;
      A.sp_n[at] = n;
      A.sp_wd[at] = A.sr_d[e];
    }
  }
}

int build_layout(const vl_window_desc* w, int n, Layout& lay, std::string& err) {
  if (n <= 0 || !w) { err = "no windows"; return 1; }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
  lay.wins.resize(n);
  lay.n_problems = 0;
  lay.data_begin = 0;
  lay.arena_bytes = 0;
  size_t max_nl = 0, max_exp_h = 0, max_exp_z = 0, max_z = 0, max_h = 0, max_gm = 0;
  int max_htiles = 0;
  WindowAnalysis plan;
  for (int i = 0; i < n; ++i) {
    const vl_window_desc& d = w[i];
    WindowLayout& W = lay.wins[i];
    char buf[256];
    if (d.mode != VL_MODE_UQS && d.mode != VL_MODE_LEP) { snprintf(buf, sizeof buf, "window %d: bad mode %d", i, d.mode); err = buf; return 1; }
    if (d.n_reads <= 0 || d.length <= 0 || d.n_clusters <= 0 || d.n_starts <= 0 || d.max_iter <= 0) {
      snprintf(buf, sizeof buf, "window %d: n_reads, length, n_clusters, n_starts and max_iter must be positive", i);
      err = buf; return 1;
    }
    if (d.n_clusters > VL_MAX_CLUSTERS) { snprintf(buf, sizeof buf, "window %d: n_clusters=%d exceeds VL_MAX_CLUSTERS=%d", i, d.n_clusters, VL_MAX_CLUSTERS); err = buf; return 1; }
    if (!d.codes) { snprintf(buf, sizeof buf, "window %d: codes is NULL (the read codes size the minority lists)", i); err = buf; return 1; }
    W.mode = d.mode; W.N = d.n_reads; W.L = d.length; W.K = d.n_clusters; W.R = d.n_starts;
    W.max_iter = d.max_iter; W.alpha0 = d.alpha0; W.conv_thres = d.conv_thres;
    W.Npad = round_up(W.N, VL_ZT_READS);
    W.Ls = round_up(W.L, VL_PT);
    W.KT = (W.K + 7) / 8; W.KS = vl_ks(W.KT); W.K8 = 8 * W.KT;
    W.n_ztiles = W.Npad / VL_ZT_READS;          // (re-derived below once the split of the haplotype tiles is known)
    plan_columns(d, W.Ls, plan);
    W.NCs = plan.NCs;
    W.n_htiles = W.NCs / VL_PT;
    // Large windows: the reads of a haplotype tile are shared by several CTAs (a property of the
    // window alone, so the sums a restart computes never depend on the rest of the batch)
    W.n_hsplit = std::max(1, std::min(VL_MAX_HSPLIT, W.Npad / VL_HSPLIT_READS));
    // Mid-sized windows (amplicon: 2 500 reads -> 39 tiles of 64 reads on 148 SMs): read tiles of 32 reads, so
    // that a lone restart's responsibility update spreads over twice the SMs.  A function of the window alone.
    W.zt_reads = (W.n_hsplit > 1 && W.Npad <= VL_FINE_ZT_MAX_READS) ? 32 : VL_ZT_READS;
    W.n_ztiles = W.Npad / W.zt_reads;
    W.nnz_cap = std::max<size_t>(plan.nnz, 1);
    W.dm = take((size_t)W.Npad * W.NCs * 8);
    W.w = take((size_t)W.Npad * 8);
    W.ltsum = take((size_t)W.Npad * 8);
    W.mcount = take((size_t)W.Npad * 8);
    W.ref = take((size_t)W.Ls);
    W.col_lb = take((size_t)W.NCs * 4);
    W.poscol = take((size_t)W.Ls * VL_N_BASES * 4);
    W.tile_pos = take((size_t)W.NCs * 4);
    W.sr_ptr = take(((size_t)W.Npad + 1) * 4);
    W.sr_idx = take(W.nnz_cap * 4);
    W.sr_d = take(W.nnz_cap * 8);
    W.sp_ptr = take(((size_t)W.Ls * VL_N_BASES + 1) * 4);
    W.sp_n = take(W.nnz_cap * 4);
    W.sp_wd = take(W.nnz_cap * 8);
    W.meta_end = off;
    W.arena_off = lay.arena_bytes;
    lay.arena_bytes += W.meta_end - W.ref;
    W.rs.resize(W.R);
    W.first_problem = lay.n_problems;
    for (int r = 0; r < W.R; ++r) {
      RestartLayout& RL = W.rs[r];
      RL.z = take((size_t)W.Npad * W.KS * 8);
      RL.wz = take((size_t)W.Npad * W.KS * 8);
      RL.h = take((size_t)W.Ls * VL_N_BASES * W.KS * 8);
      RL.g = take((size_t)W.Ls * VL_N_BASES * W.KS * 8);
      RL.gm = take((size_t)W.NCs * W.KS * 8);
      RL.zpart = take((size_t)W.n_ztiles * (W.K8 + VL_ZP_EXTRA) * 8);
      RL.hpart = take((size_t)W.n_htiles * VL_HP_STRIDE * 8);
      RL.tsplit = RL.hsplit_count = 0;
      if (W.n_hsplit > 1) {
        RL.tsplit = take((size_t)W.n_htiles * W.n_hsplit * VL_PT * W.K8 * 8);
        RL.hsplit_count = take((size_t)W.n_htiles * 4);
      }
      RL.alpha = take((size_t)W.K8 * 8);
      RL.mlp = take((size_t)W.K8 * 8);
      RL.mlp_lay = take((size_t)W.K8 * 8);
      RL.lay = take((size_t)W.K8 * 4);
      RL.gat = take((size_t)W.K8 * 4);
      RL.lay_z = take((size_t)W.K8 * 4);
    }
    // history_elbo of the R restarts, contiguous so that one 2-D copy fetches them
    const size_t hist0 = take((size_t)W.R * W.max_iter * 8);
    for (int r = 0; r < W.R; ++r) W.rs[r].hist = hist0 + (size_t)r * W.max_iter * 8;
    lay.n_problems += W.R;
    max_nl = std::max(max_nl, (size_t)W.N * W.L);
    max_exp_h = std::max(max_exp_h, (size_t)W.K * W.L * VL_N_BASES * 8);
    max_exp_z = std::max(max_exp_z, (size_t)W.N * W.K * 8);
    max_z = std::max(max_z, (size_t)W.Npad * W.KS * 8);
    max_h = std::max(max_h, (size_t)W.Ls * VL_N_BASES * W.KS * 8);
    max_gm = std::max(max_gm, (size_t)W.NCs * W.KS * 8);
    max_htiles = std::max(max_htiles, W.n_htiles * W.n_hsplit);
    lay.n_htiles_total += (size_t)W.n_htiles * W.n_hsplit * W.R;
    lay.n_ztiles_total += (size_t)W.n_ztiles * W.R;
  }
  lay.data_end = off;
  lay.probs = take((size_t)lay.n_problems * sizeof(VlProblem));
  lay.states = take((size_t)lay.n_problems * sizeof(VlState));
  lay.done = take(64);
  lay.pstat = take((size_t)lay.n_problems * sizeof(int4));
  lay.htiles = take(lay.n_htiles_total * sizeof(int2));
  lay.ztiles = take(lay.n_ztiles_total * sizeof(int2));
  lay.rlist = take((size_t)lay.n_problems * sizeof(int));
  lay.cbase = take((size_t)lay.n_problems * sizeof(int4));
  lay.ctl = take((size_t)lay.n_problems * sizeof(VlCtl));
  lay.nlive = take(VL_N_CLASSES * 4 * sizeof(int));
  lay.raw_codes = take(max_nl);
  lay.raw_lt = take(max_nl * 8);
  lay.raw_lo = take(max_nl * 8);
  lay.export_h = take(max_exp_h);
  lay.export_z = take(max_exp_z);
  lay.tmp_z = take(max_z);
  lay.tmp_wz = take(max_z);
  lay.tmp_h = take(max_h);
  lay.tmp_g = take(max_h);
  lay.tmp_gm = take(max_gm);
  lay.tmp_hpart = take((size_t)max_htiles * VL_HP_STRIDE * 8);
  lay.tmp_group = take(VL_MAX_CLUSTERS * sizeof(int));
  lay.tmp_prob = take(sizeof(VlProblem));
  lay.tmp_state = take(sizeof(VlState));
  lay.tmp_tiles = take((size_t)max_htiles * sizeof(int2));
  lay.tmp_lay = take(VL_MAX_CLUSTERS * sizeof(int));
  lay.tmp_gat = take(VL_MAX_CLUSTERS * sizeof(int));
  lay.tmp_mlp = take(VL_MAX_CLUSTERS * sizeof(double));
  lay.total = off;
  return 0;
}

}  // namespace

struct vl_batch {
  int device = 0;
  cudaStream_t stream = nullptr;
  unsigned char* ws = nullptr;
  size_t ws_bytes = 0;
  Layout lay;
  std::vector<VlProblem> probs;        // host mirror of the device problem table
  std::vector<int> order;              // problems, costliest window first
  std::vector<VlState> states;         // host mirror filled by fetch / upload
  std::vector<int4> pstat;             // host mirror: (running, tiles needed, valid slots, updates)
  // pinned staging of the per-chunk schedule
  int2* h_htiles = nullptr; int2* h_ztiles = nullptr; int4* h_pstat = nullptr;
  int* h_rlist = nullptr;
  int4* h_cbase = nullptr;            // per problem: updates done when the chunk was scheduled, haplotype tiles (x parts), read tiles
  int* h_nlive = nullptr;
  unsigned char* h_arena = nullptr;    // pinned image of every window's small arrays (maps, sparse lists)
  // tile lists (per-phase launches) and restart lists (chunk launches) per kernel class and variant
  // v = mode (bit 0: lep) + 2 * (window with split haplotype tiles)
  int n_ht[VL_N_CLASSES][4] = {}, n_zt[VL_N_CLASSES][4] = {}, n_rs[VL_N_CLASSES][4] = {};
  size_t off_ht[VL_N_CLASSES][4] = {}, off_zt[VL_N_CLASSES][4] = {}, off_rs[VL_N_CLASSES][4] = {};
  int kt_max[VL_N_CLASSES][4] = {};         // widest layout among the restarts of a list (cluster tiles)
  int want_workers[VL_N_CLASSES][4] = {};   // tiles the restarts of a list can have in flight at once
  int n_sched = 0;
  bool sched_dirty = true;
  int skip_dead = 1, fused = 1;
  cudaStream_t class_stream[VL_N_CLASSES * 4] = {};   // chunk launches of different classes run side by side
  cudaEvent_t ev_fork = nullptr, ev_join[VL_N_CLASSES * 4] = {};
  int occ[VL_N_CLASSES][4][3] = {};    // resident CTAs per SM of the persistent chunk kernel (class, variant, deep), 0 = not asked yet
  int persistent = -1;                 // chunk launches: 1 = persistent workers, 0 = one block per tile, -1 = by the number of launches
  int n_sms = 148;
  int sm_priority = 1;                 // SM shares of side-by-side launches: widest class first (see run_chunk_fused)
  int merge_classes = 1;               // one launch of the widest class needed per variant (see schedule)
  int deep2_max = 296;                 // most tiles in flight for which a launch gives every CTA a whole SM's shared memory
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_ms = 0.0;
  bool uploaded = false;

  template <typename T> T* dev(size_t off) const { return reinterpret_cast<T*>(ws + off); }
  VlProblem* d_probs() const { return dev<VlProblem>(lay.probs); }
  VlState* d_states() const { return dev<VlState>(lay.states); }
  int* d_done() const { return dev<int>(lay.done); }
  int4* d_pstat() const { return dev<int4>(lay.pstat); }
};

namespace {

// Where the still-running restarts go next: grouped by kernel class (layout width), mode and
// whether the window's haplotype tiles are split.  Tile lists serve the per-phase launches, restart
// lists the chunk launches.
int schedule(vl_batch* b) {
  const Layout& L = b->lay;
  std::vector<int2> ht[VL_N_CLASSES][4], zt[VL_N_CLASSES][4];
  std::vector<int> rs[VL_N_CLASSES][4];
  b->n_sched = 0;
  for (int c = 0; c < VL_N_CLASSES; ++c)
    for (int m = 0; m < 4; ++m) { b->want_workers[c][m] = 0; b->kt_max[c][m] = 0; }
  // The restarts of a variant (mode, split haplotype tiles) all go to the widest kernel class any of them
  // needs, so that a chunk is ONE launch of one-tile blocks per variant.  Launching every class on its own
  // -- persistent workers side by side, each with a static share of the SMs -- was measured slower on every
  // workload (hiv5k 647 -> 592 ms, amplicon 2.92 -> 2.67 s): a wide straggler kept only its share of the
  // SMs, and narrow tiles lose little in the wider class's kernel.  The arithmetic of a restart is the
  // same in every class that holds its layout (test: ragged batch == single-window runs, bitwise).
  // merge_classes = 0 keeps every class apart (development aid).
  int top_class[4] = {-1, -1, -1, -1};
  if (b->merge_classes)
    for (int p : b->order) {
      const int4 st = b->pstat[p];
      if (st.x == 0) continue;
      const VlProblem& P = b->probs[p];
      const int v = (P.mode == VL_MODE_LEP ? 1 : 0) + (P.n_hsplit > 1 ? 2 : 0);
      top_class[v] = std::max(top_class[v], vl_class_of_tiles(st.y));
    }
  for (int p : b->order) {
    const int4 st = b->pstat[p];
    if (st.x == 0) continue;
    const VlProblem& P = b->probs[p];
    const int m = P.mode == VL_MODE_LEP ? 1 : 0;
    const int v = m + (P.n_hsplit > 1 ? 2 : 0);
    const int cls = top_class[v] >= 0 ? top_class[v] : vl_class_of_tiles(st.y);
    for (int sp = 0; sp < P.n_hsplit; ++sp)
      for (int t = 0; t < P.n_htiles; ++t) ht[cls][v].push_back(make_int2(p, vl_htile_id(t, sp, P.n_hsplit)));
    for (int t = 0; t < P.n_ztiles; ++t) zt[cls][v].push_back(make_int2(p, t));
    rs[cls][v].push_back(p);
    b->want_workers[cls][v] += std::max(P.n_htiles * P.n_hsplit, P.n_ztiles);
    b->kt_max[cls][v] = std::max(b->kt_max[cls][v], st.y);
    ++b->n_sched;
  }
  size_t ho = 0, zo = 0, ro = 0;
  for (int c = 0; c < VL_N_CLASSES; ++c)
    for (int m = 0; m < 4; ++m) {
      b->n_ht[c][m] = (int)ht[c][m].size(); b->off_ht[c][m] = ho;
      b->n_zt[c][m] = (int)zt[c][m].size(); b->off_zt[c][m] = zo;
      b->n_rs[c][m] = (int)rs[c][m].size(); b->off_rs[c][m] = ro;
      std::copy(ht[c][m].begin(), ht[c][m].end(), b->h_htiles + ho);
      std::copy(zt[c][m].begin(), zt[c][m].end(), b->h_ztiles + zo);
      std::copy(rs[c][m].begin(), rs[c][m].end(), b->h_rlist + ro);
      ho += ht[c][m].size(); zo += zt[c][m].size(); ro += rs[c][m].size();
    }
  cudaStream_t s = b->stream;
  if (ho) VL_CUDA(cudaMemcpyAsync(b->dev<int2>(L.htiles), b->h_htiles, ho * sizeof(int2), cudaMemcpyHostToDevice, s));
  if (zo) VL_CUDA(cudaMemcpyAsync(b->dev<int2>(L.ztiles), b->h_ztiles, zo * sizeof(int2), cudaMemcpyHostToDevice, s));
  if (ro) VL_CUDA(cudaMemcpyAsync(b->dev<int>(L.rlist), b->h_rlist, ro * sizeof(int), cudaMemcpyHostToDevice, s));
  b->sched_dirty = false;
  return 0;
}

int launch_phase(vl_batch* b, int phase, int64_t* launches) {
  const Layout& L = b->lay;
  for (int c = 0; c < VL_N_CLASSES; ++c)
    for (int m = 0; m < 4; ++m) {
      const int mode = (m & 1) ? VL_MODE_LEP : VL_MODE_UQS;
      if (phase == 0 && b->n_ht[c][m] > 0) {
        VL_CUDA(vl_launch_haplo(c, mode, m >> 1, b->d_probs(), b->dev<int2>(L.htiles) + b->off_ht[c][m], b->n_ht[c][m], b->stream));
        *launches += 1;
      }
      if (phase == 1 && b->n_zt[c][m] > 0) {
        VL_CUDA(vl_launch_cluster(c, mode, b->d_probs(), b->dev<int2>(L.ztiles) + b->off_zt[c][m], b->n_zt[c][m],
                                  b->d_done(), b->d_pstat(), b->stream));
        *launches += 1;
      }
    }
  return 0;
}

// n_iter updates of every scheduled restart: one launch per kernel class and variant (vl_sched_kernel),
// the launches side by side on their own streams.  One launch alone: one block per tile the chunk can
// have.  Several launches: persistent workers, each launch as many as its restarts can have tiles in
// flight, at most what fits the GPU; when the launches together want more SMs than there are, each gets
// its share (blocks of a huge one-tile-per-block grid would fill every SM slot and keep the other
// launches out).
int run_chunk_fused(vl_batch* b, int n_iter, int64_t* launches) {
  const Layout& L = b->lay;
  const int np = L.n_problems;
  for (int p = 0; p < np; ++p)
    b->h_cbase[p] = make_int4(b->pstat[p].w, b->probs[p].n_htiles * b->probs[p].n_hsplit, b->probs[p].n_ztiles, 0);
  VL_CUDA(cudaMemcpyAsync(b->dev<int4>(L.cbase), b->h_cbase, (size_t)np * sizeof(int4), cudaMemcpyHostToDevice, b->stream));
  VL_CUDA(cudaMemsetAsync(b->dev<VlCtl>(L.ctl), 0, (size_t)np * sizeof(VlCtl), b->stream));
  struct Launch { int c, m, deep, occ, workers; double sms; };
  std::vector<Launch> todo;
  for (int c = 0; c < VL_N_CLASSES; ++c)
    for (int m = 0; m < 4; ++m) {
      b->h_nlive[c * 4 + m] = b->n_rs[c][m];
      if (b->n_rs[c][m] > 0) todo.push_back(Launch{c, m, 0, 0, 0, 0.0});
    }
  if (todo.empty()) return 0;
  // Few tiles in flight: the SMs are mostly idle and a restart's update is a chain of latencies, so a
  // CTA gets more shared memory (larger pipeline stages): a whole SM's when all launches together have
  // at most one tile per SM in flight, half an SM's when the launch has at most two per SM.
  {
    long long want_all = 0;
    for (const Launch& l : todo) want_all += b->want_workers[l.c][l.m];
    for (Launch& l : todo) l.deep = want_all <= b->deep2_max ? 2 : (b->want_workers[l.c][l.m] <= 2 * b->n_sms ? 1 : 0);
  }
  VL_CUDA(cudaMemcpyAsync(b->dev<int>(L.nlive), b->h_nlive, VL_N_CLASSES * 4 * sizeof(int), cudaMemcpyHostToDevice, b->stream));
  const bool fork = todo.size() > 1;
  const int persistent = b->persistent >= 0 ? b->persistent : (fork ? 1 : 0);
  double sm_demand = 0.0;
  if (persistent)
    for (Launch& l : todo) {
      int& occ = b->occ[l.c][l.m][l.deep];
      if (occ == 0) {
        int smem = 0, regs = 0;
        VL_CUDA(vl_sched_occupancy(l.c, (l.m & 1) ? VL_MODE_LEP : VL_MODE_UQS, l.m >> 1, l.deep, 1, &occ, &smem, &regs));
        if (occ <= 0) return fail("the chunk kernel does not fit an SM");
      }
      l.occ = occ;
      l.workers = std::min(b->want_workers[l.c][l.m], occ * b->n_sms);
      l.sms = (double)l.workers / occ;
      sm_demand += l.sms;
    }
  // More demand than SMs: a chunk ends when its slowest launch has done its updates, and a launch needs
  // (tiles in flight / workers) rounds of its restarts' update time per update -- about 32 + 8 kt us for
  // a layout of kt cluster tiles (measured on lone restarts).  Equal finishing times: SM shares in
  // proportion to demand x update time (demand alone halves the workers of the wide restarts a run ends
  // up waiting for); what a launch cannot use goes to the others.
  double scale = sm_demand > b->n_sms ? b->n_sms / sm_demand : 1.0;
  if (persistent && sm_demand > b->n_sms && b->sm_priority) {
    double wsum = 0.0;
    std::vector<double> w(todo.size());
    for (size_t k = 0; k < todo.size(); ++k) { w[k] = todo[k].sms * (32.0 + 8.0 * b->kt_max[todo[k].c][todo[k].m]); wsum += w[k]; }
    double left = b->n_sms, wleft = wsum;
    std::vector<char> fixed(todo.size(), 0);
    for (int pass = 0; pass < 2; ++pass)                 // launches whose share exceeds their demand keep the demand
      for (size_t k = 0; k < todo.size(); ++k)
        if (!fixed[k] && left * w[k] / wleft >= todo[k].sms) { fixed[k] = 1; left -= todo[k].sms; wleft -= w[k]; }
    for (size_t k = 0; k < todo.size(); ++k) {
      Launch& l = todo[k];
      const double grant = fixed[k] ? l.sms : std::max(1.0, left * w[k] / wleft);
      l.workers = std::max(1, std::min(l.workers, (int)(grant * l.occ + 0.5)));
      l.sms = grant;
    }
    scale = 1.0;
  }
  if (fork) VL_CUDA(cudaEventRecord(b->ev_fork, b->stream));
  if (b->sm_priority) std::reverse(todo.begin(), todo.end());      // widest class first: its blocks are placed before the small ones fill the SMs
  for (const Launch& l : todo) {
    const int i = l.c * 4 + l.m;
    cudaStream_t s = fork ? b->class_stream[i] : b->stream;
    if (fork) VL_CUDA(cudaStreamWaitEvent(s, b->ev_fork, 0));
    const long long blocks = persistent ? std::max(1, std::min(l.workers, (int)(l.sms * scale * l.occ + 0.5)))
                                        : (long long)n_iter * (b->n_ht[l.c][l.m] + b->n_zt[l.c][l.m]);
    VL_CUDA(vl_launch_sched(l.c, (l.m & 1) ? VL_MODE_LEP : VL_MODE_UQS, l.m >> 1, b->d_probs(), b->dev<int>(L.rlist) + b->off_rs[l.c][l.m],
                            b->n_rs[l.c][l.m], n_iter, b->dev<int4>(L.cbase), b->dev<VlCtl>(L.ctl), b->dev<int>(L.nlive) + i,
                            b->d_done(), b->d_pstat(), blocks, l.deep, persistent, s));
    *launches += 1;
    if (fork) {
      VL_CUDA(cudaEventRecord(b->ev_join[i], s));
      VL_CUDA(cudaStreamWaitEvent(b->stream, b->ev_join[i], 0));
    }
  }
  return 0;
}

int run_iterations(vl_batch* b, int n_iter, int64_t* launches) {
  for (int it = 0; it < n_iter; ++it)
    for (int phase = 0; phase < 2; ++phase)
      if (launch_phase(b, phase, launches)) return 1;
  return 0;
}

// Read the per-restart status back and note whether the schedule has to be rebuilt.
int sync_status(vl_batch* b, int* n_running) {
  const int np = b->lay.n_problems;
  VL_CUDA(cudaMemcpyAsync(b->h_pstat, b->d_pstat(), (size_t)np * sizeof(int4), cudaMemcpyDeviceToHost, b->stream));
  VL_CUDA(cudaStreamSynchronize(b->stream));
  int running = 0;
  for (int p = 0; p < np; ++p) {
    const int4 now = b->h_pstat[p], old = b->pstat[p];
    if (now.x != old.x || now.y != old.y) b->sched_dirty = true;
    b->pstat[p] = now;
    running += now.x;
  }
  if (n_running) *n_running = running;
  return 0;
}

void identity_layout(int K, int K8, std::vector<int>& lay, std::vector<int>& gat) {
  lay.assign(K8, -1); gat.assign(K8, 0);
  for (int k = 0; k < K; ++k) { lay[k] = k; gat[k] = k; }
}

}  // namespace

// for the other translation units of the library (declared in vl_launch.h)
int vl_set_error(const char* msg) { return fail(msg); }

extern "C" {
#pragma GCC visibility push(default)

int vl_abi_version(void) { return VL_ABI_VERSION; }

const char* vl_last_error(void) { return g_err.c_str(); }

int vl_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return -1; }
  return n;
}

int vl_batch_workspace_bytes(const vl_window_desc* windows, int32_t n_windows, size_t* bytes_out) {
  if (!bytes_out) return fail("bytes_out is NULL");
  Layout lay;
  std::string err;
  if (build_layout(windows, n_windows, lay, err)) return fail(err);
  *bytes_out = lay.total;
  return 0;
}

int32_t vl_window_plan_capacity(int32_t length) {
  return length <= 0 ? 0 : VL_PT * ((5 * length + 11) / 12 + 1);
}

int vl_window_plan(const vl_window_desc* w, int32_t* n_cols_out, int64_t* n_sparse_out, int32_t* col_lb) {
  if (!w || !w->codes) return fail("window or its codes is NULL");
  if (w->n_reads <= 0 || w->length <= 0) return fail("n_reads and length must be positive");
  WindowAnalysis A;
  plan_columns(*w, round_up(w->length, VL_PT), A);
  if (n_cols_out) *n_cols_out = A.NCs;
  if (n_sparse_out) *n_sparse_out = (int64_t)A.nnz;
  if (col_lb) std::copy(A.col_lb.begin(), A.col_lb.end(), col_lb);
  return 0;
}

int vl_batch_create(vl_batch** out, int32_t device, const vl_window_desc* windows, int32_t n_windows,
                    void* dev_workspace, size_t dev_bytes, void* stream) {
  if (!out) return fail("out is NULL");
  *out = nullptr;
  if (!dev_workspace) return fail("dev_workspace is NULL");
  if ((reinterpret_cast<uintptr_t>(dev_workspace) & (kAlign - 1)) != 0) return fail("dev_workspace must be 256-byte aligned");
  vl_batch* b = new vl_batch();
  std::string err;
  if (build_layout(windows, n_windows, b->lay, err)) { delete b; return fail(err); }
  if (dev_bytes < b->lay.total) {
    char buf[160];
    snprintf(buf, sizeof buf, "workspace too small: %zu bytes given, %zu needed", dev_bytes, b->lay.total);
    delete b;
    return fail(buf);
  }
  b->device = device;
  b->stream = reinterpret_cast<cudaStream_t>(stream);
  b->ws = static_cast<unsigned char*>(dev_workspace);
  b->ws_bytes = dev_bytes;
  if (const char* env = std::getenv("VL_PERSISTENT")) b->persistent = std::atoi(env);      // development aid
  if (const char* env = std::getenv("VL_DEEP2_MAX")) b->deep2_max = std::atoi(env);        // development aid
  if (const char* env = std::getenv("VL_MERGE_CLASSES")) b->merge_classes = std::atoi(env); // development aid
  if (const char* env = std::getenv("VL_SM_PRIORITY")) b->sm_priority = std::atoi(env);     // development aid
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete b; return fail(std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }

  const Layout& L = b->lay;
  b->probs.resize(L.n_problems);
  b->states.resize(L.n_problems);
  b->pstat.assign(L.n_problems, make_int4(0, 0, 0, 0));
  // within a kernel class the costliest windows come first so that the long CTAs start early
  std::vector<int> worder(L.wins.size());
  for (size_t i = 0; i < worder.size(); ++i) worder[i] = (int)i;
  std::stable_sort(worder.begin(), worder.end(), [&](int x, int y) {
    const WindowLayout& A = L.wins[x]; const WindowLayout& B = L.wins[y];
    return (double)A.Npad * A.Ls > (double)B.Npad * B.Ls;
  });
  for (size_t i = 0; i < L.wins.size(); ++i) {
    const WindowLayout& W = L.wins[i];
    for (int r = 0; r < W.R; ++r) {
      VlProblem& P = b->probs[W.first_problem + r];
      P.mode = W.mode; P.N = W.N; P.L = W.L; P.K = W.K;
      P.Npad = W.Npad; P.Ls = W.Ls; P.KT = W.KT; P.NCs = W.NCs;
      P.n_ztiles = W.n_ztiles; P.n_htiles = W.n_htiles; P.window = (int)i; P.restart = r;
      P.n_hsplit = W.n_hsplit; P.zt_reads = W.zt_reads;
      P.dm = b->dev<double>(W.dm);
      P.w = b->dev<double>(W.w);
      P.ltsum = b->dev<double>(W.ltsum);
      P.mcount = b->dev<double>(W.mcount);
      P.ref = b->dev<uint8_t>(W.ref);
      P.col_lb = b->dev<int>(W.col_lb); P.poscol = b->dev<int>(W.poscol); P.tile_pos = b->dev<int>(W.tile_pos);
      P.sr_ptr = b->dev<int>(W.sr_ptr); P.sr_idx = b->dev<int>(W.sr_idx); P.sr_d = b->dev<double>(W.sr_d);
      P.sp_ptr = b->dev<int>(W.sp_ptr); P.sp_n = b->dev<int>(W.sp_n); P.sp_d = b->dev<double>(W.sp_wd);
      const RestartLayout& RL = W.rs[r];
      P.z = b->dev<double>(RL.z); P.wz = b->dev<double>(RL.wz); P.h = b->dev<double>(RL.h);
      P.g = b->dev<double>(RL.g); P.gm = b->dev<double>(RL.gm);
      P.zpart = b->dev<double>(RL.zpart); P.hpart = b->dev<double>(RL.hpart);
      P.tsplit = W.n_hsplit > 1 ? b->dev<double>(RL.tsplit) : nullptr;
      P.hsplit_count = W.n_hsplit > 1 ? b->dev<int>(RL.hsplit_count) : nullptr;
      P.alpha = b->dev<double>(RL.alpha); P.mlp = b->dev<double>(RL.mlp);
      P.mlp_lay = b->dev<double>(RL.mlp_lay);
      P.lay = b->dev<int>(RL.lay); P.gat = b->dev<int>(RL.gat); P.lay_z = b->dev<int>(RL.lay_z);
      P.hist = b->dev<double>(RL.hist);
      P.st = b->d_states() + (W.first_problem + r);
    }
  }
  for (int wi : worder)
    for (int r = 0; r < L.wins[wi].R; ++r) b->order.push_back(L.wins[wi].first_problem + r);

#define VL_CREATE_CUDA(expr)                                                                    \
  do {                                                                                          \
    cudaError_t e2_ = (expr);                                                                   \
    if (e2_ != cudaSuccess) {                                                                   \
      std::string m_ = std::string("CUDA error in vl_batch_create: ") + cudaGetErrorString(e2_); \
      vl_batch_destroy(b);                                                                      \
      return fail(m_);                                                                          \
    }                                                                                           \
  } while (0)
  VL_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&b->h_htiles), std::max<size_t>(L.n_htiles_total, 1) * sizeof(int2)));
  VL_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&b->h_ztiles), std::max<size_t>(L.n_ztiles_total, 1) * sizeof(int2)));
  VL_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&b->h_pstat), (size_t)L.n_problems * sizeof(int4)));
  VL_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&b->h_rlist), (size_t)L.n_problems * sizeof(int)));
  VL_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&b->h_cbase), (size_t)L.n_problems * sizeof(int4)));
  VL_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&b->h_nlive), VL_N_CLASSES * 4 * sizeof(int)));
  for (int i = 0; i < VL_N_CLASSES * 4; ++i) {
    VL_CREATE_CUDA(cudaStreamCreateWithFlags(&b->class_stream[i], cudaStreamNonBlocking));
    VL_CREATE_CUDA(cudaEventCreateWithFlags(&b->ev_join[i], cudaEventDisableTiming));
  }
  VL_CREATE_CUDA(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
  VL_CREATE_CUDA(cudaDeviceGetAttribute(&b->n_sms, cudaDevAttrMultiProcessorCount, device));
  VL_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&b->h_arena), std::max<size_t>(L.arena_bytes, 1)));
  VL_CREATE_CUDA(cudaEventCreate(&b->ev0));
  VL_CREATE_CUDA(cudaEventCreate(&b->ev1));
  VL_CREATE_CUDA(cudaMemcpyAsync(b->d_probs(), b->probs.data(), b->probs.size() * sizeof(VlProblem),
                                 cudaMemcpyHostToDevice, b->stream));
  VL_CREATE_CUDA(cudaStreamSynchronize(b->stream));
#undef VL_CREATE_CUDA
  *out = b;
  return 0;
}

void vl_batch_destroy(vl_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->h_htiles) cudaFreeHost(b->h_htiles);
  if (b->h_ztiles) cudaFreeHost(b->h_ztiles);
  if (b->h_pstat) cudaFreeHost(b->h_pstat);
  if (b->h_rlist) cudaFreeHost(b->h_rlist);
  if (b->h_cbase) cudaFreeHost(b->h_cbase);
  if (b->h_nlive) cudaFreeHost(b->h_nlive);
  for (int i = 0; i < VL_N_CLASSES * 4; ++i) {
    if (b->class_stream[i]) cudaStreamDestroy(b->class_stream[i]);
    if (b->ev_join[i]) cudaEventDestroy(b->ev_join[i]);
  }
  if (b->ev_fork) cudaEventDestroy(b->ev_fork);
  if (b->h_arena) cudaFreeHost(b->h_arena);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  delete b;
}

int vl_batch_set_option(vl_batch* b, int32_t option, int32_t value) {
  if (!b) return fail("batch is NULL");
  if (option == VL_OPT_SKIP_DEAD) { b->skip_dead = value ? 1 : 0; return 0; }
  if (option == VL_OPT_FUSED_CHUNKS) { b->fused = value ? 1 : 0; return 0; }
  return fail("unknown option");
}

int vl_batch_upload(vl_batch* b, const vl_window_desc* windows, int32_t n_windows) {
  if (!b) return fail("batch is NULL");
  const Layout& L = b->lay;
  if (n_windows != (int)L.wins.size()) return fail("n_windows differs from the batch");
  VL_CUDA(cudaSetDevice(b->device));
  cudaStream_t s = b->stream;
  static const bool trace = std::getenv("VL_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(s);
    fprintf(stderr, "upload: %-28s at %.2f ms\n", what,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
  };
  VL_CUDA(cudaMemsetAsync(b->ws + L.data_begin, 0, L.data_end - L.data_begin, s));
  lap("workspace cleared");
  for (int i = 0; i < n_windows; ++i) {
    const vl_window_desc& d = windows[i];
    const WindowLayout& W = L.wins[i];
    if (d.mode != W.mode || d.n_reads != W.N || d.length != W.L || d.n_clusters != W.K || d.n_starts != W.R)
      return fail("window shape differs from the one the batch was created with");
    if (!d.codes || !d.weights || !d.ref_codes || !d.init_cluster || !d.init_scalars)
      return fail("window input pointer is NULL");
    if (W.mode == VL_MODE_UQS && (!d.log_hit || !d.log_off)) return fail("uqs window without log_hit/log_off");
  }
  // window analysis (dense columns, sparse lists) on host threads, written into the pinned arena
  // in the device layout of the block [ref .. sp_wd] so that one copy per window uploads it
  std::vector<int> bad(n_windows, 0);
  {
    const int n_threads = std::max(1, std::min<int>({(int)std::thread::hardware_concurrency(), 16, n_windows}));
    auto work = [&](int t) {
      WindowAnalysis A;
      for (int i = t; i < n_windows; i += n_threads) {
        const vl_window_desc& d = windows[i];
        const WindowLayout& W = L.wins[i];
        analyze_window(d, W.Npad, W.Ls, A);
        if (A.nnz > W.nnz_cap || A.NCs != W.NCs) { bad[i] = 1; continue; }
        unsigned char* base = b->h_arena + W.arena_off;
        auto at = [&](size_t ws_off) { return base + (ws_off - W.ref); };
        std::memset(base, 0, W.meta_end - W.ref);
        std::memset(at(W.ref), VL_CODE_N, W.Ls);
        std::memcpy(at(W.ref), d.ref_codes, W.L);
        std::memcpy(at(W.col_lb), A.col_lb.data(), (size_t)W.NCs * 4);
        std::memcpy(at(W.poscol), A.poscol.data(), A.poscol.size() * 4);
        std::memcpy(at(W.tile_pos), A.tile_pos.data(), (size_t)W.NCs * 4);
        std::memcpy(at(W.sr_ptr), A.sr_ptr.data(), A.sr_ptr.size() * 4);
        std::memcpy(at(W.sp_ptr), A.sp_ptr.data(), A.sp_ptr.size() * 4);
        if (A.nnz) {
          std::memcpy(at(W.sr_idx), A.sr_idx.data(), A.nnz * 4);
          std::memcpy(at(W.sr_d), A.sr_d.data(), A.nnz * 8);
          std::memcpy(at(W.sp_n), A.sp_n.data(), A.nnz * 4);
          std::memcpy(at(W.sp_wd), A.sp_wd.data(), A.nnz * 8);
        }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
  }
  lap("host analysis");
  for (int i = 0; i < n_windows; ++i)
    if (bad[i]) {
      char buf[200];
      snprintf(buf, sizeof buf, "window %d: its reads need more dense columns / sparse entries than the batch was "
               "created for; create a new batch for these reads", i);
      return fail(buf);
    }
  for (int i = 0; i < n_windows; ++i) {
    const vl_window_desc& d = windows[i];
    const WindowLayout& W = L.wins[i];
    const bool uqs = W.mode == VL_MODE_UQS;
    const size_t nl = (size_t)W.N * W.L;
    // the staging area of the raw arrays is reused window after window; stream order keeps that safe
    VL_CUDA(cudaMemcpyAsync(b->ws + L.raw_codes, d.codes, nl, cudaMemcpyHostToDevice, s));
    if (uqs) {
      VL_CUDA(cudaMemcpyAsync(b->ws + L.raw_lt, d.log_hit, nl * 8, cudaMemcpyHostToDevice, s));
      VL_CUDA(cudaMemcpyAsync(b->ws + L.raw_lo, d.log_off, nl * 8, cudaMemcpyHostToDevice, s));
    }
    VL_CUDA(cudaMemcpyAsync(b->ws + W.ref, b->h_arena + W.arena_off, W.meta_end - W.ref, cudaMemcpyHostToDevice, s));
    VL_CUDA(cudaMemcpyAsync(b->ws + W.w, d.weights, (size_t)W.N * 8, cudaMemcpyHostToDevice, s));
    VL_CUDA(vl_launch_prepare(b->dev<uint8_t>(L.raw_codes), b->dev<double>(L.raw_lt), b->dev<double>(L.raw_lo),
                              b->dev<int>(W.col_lb), b->dev<double>(W.dm), b->dev<double>(W.mcount),
                              b->dev<double>(W.ltsum), W.N, W.L, W.Npad, W.NCs, W.mode, s));
    for (int r = 0; r < W.R; ++r) {
      // one contiguous copy into the staging area (the export buffer, idle during uploads; reuse is
      // stream-ordered), the row pitch is applied on the device
      VL_CUDA(cudaMemcpyAsync(b->ws + L.export_z, d.init_cluster + (size_t)r * W.N * W.K, (size_t)W.N * W.K * 8,
                              cudaMemcpyHostToDevice, s));
      VL_CUDA(vl_launch_init_rows(b->dev<double>(L.export_z), b->dev<double>(W.w), b->dev<double>(W.rs[r].z),
                                  b->dev<double>(W.rs[r].wz), W.N, W.K, W.Npad, W.KS, s));
      VlState& S = b->states[W.first_problem + r];
      std::memset(&S, 0, sizeof S);
      const double* in = d.init_scalars + (size_t)r * VL_INIT_STRIDE;
      S.alpha0 = W.alpha0; S.thr = W.conv_thres;
      S.a0 = in[0]; S.b0 = in[1]; S.mlg0 = in[2]; S.mlg1 = in[3];
      S.c0 = in[4]; S.d0 = in[5]; S.mlt0 = in[6]; S.mlt1 = in[7];
      S.max_iter = W.max_iter;
      S.skip = b->skip_dead;
    }
  }
  lap("copies + operand kernels");
  VL_CUDA(cudaMemcpyAsync(b->d_states(), b->states.data(), b->states.size() * sizeof(VlState), cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemsetAsync(b->d_done(), 0, 64, s));
  VL_CUDA(vl_launch_init(b->d_probs(), L.n_problems, b->d_pstat(), s));
  if (sync_status(b, nullptr)) return 1;
  b->sched_dirty = true;
  b->uploaded = true;
  return 0;
}

int vl_batch_run(vl_batch* b, int64_t* n_launches_out) {
  if (!b) return fail("batch is NULL");
  if (!b->uploaded) return fail("vl_batch_run before vl_batch_upload");
  VL_CUDA(cudaSetDevice(b->device));
  int64_t launches = 0;
  int max_iter = 0;
  for (const WindowLayout& W : b->lay.wins) max_iter = std::max(max_iter, W.max_iter);
  VL_CUDA(cudaEventRecord(b->ev0, b->stream));
  int running = 1, n_chunks = 0;
  for (int64_t done_iters = 0; running > 0 && done_iters < (int64_t)max_iter + 256; ++n_chunks) {
    // the layouts shrink fastest in the first iterations: short chunks there keep the restarts
    // in the narrowest kernel class that fits them
    // (a fixed function of the chunk index: the path a restart takes never depends on the
    // other windows of the batch, so results are independent of the batch composition)
    const int chunk = n_chunks < 6 ? 4 : (n_chunks < 14 ? 32 : (n_chunks < 46 ? 64 : 256));
    static const bool trace = std::getenv("VL_TRACE") != nullptr;    // development aid: chunk timeline on stderr
    const auto t0 = std::chrono::steady_clock::now();
    const int before = running;
    if (b->sched_dirty && schedule(b)) return 1;
    if (b->fused ? run_chunk_fused(b, chunk, &launches) : run_iterations(b, chunk, &launches)) return 1;
    if (sync_status(b, &running)) return 1;
    done_iters += chunk;
    if (trace) {
      const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
      int kt_min = 1 << 30, kt_max = 0;
      for (const int4& st : b->pstat)
        if (st.x) { kt_min = std::min(kt_min, st.y); kt_max = std::max(kt_max, st.y); }
      fprintf(stderr, "chunk %3d: %3d updates, running %d -> %d, %.0f us (%.1f us/update), cluster tiles of the running %d..%d\n",
              n_chunks, chunk, before, running, us, us / chunk, running ? kt_min : 0, kt_max);
    }
  }
  VL_CUDA(cudaEventRecord(b->ev1, b->stream));
  VL_CUDA(cudaEventSynchronize(b->ev1));
  float ms = 0.f;
  VL_CUDA(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  b->last_ms = ms;
  if (n_launches_out) *n_launches_out = launches;
  return 0;
}

int vl_batch_step(vl_batch* b, int32_t n_updates, int64_t* n_launches_out) {
  if (!b) return fail("batch is NULL");
  if (!b->uploaded) return fail("vl_batch_step before vl_batch_upload");
  if (n_updates < 0) return fail("n_updates < 0");
  VL_CUDA(cudaSetDevice(b->device));
  int64_t launches = 0;
  VL_CUDA(cudaEventRecord(b->ev0, b->stream));
  // one update per schedule: a layout may move to another kernel class at any update
  for (int it = 0; it < n_updates; ++it) {
    if (b->sched_dirty && schedule(b)) return 1;
    if (run_iterations(b, 1, &launches)) return 1;
    if (sync_status(b, nullptr)) return 1;
  }
  VL_CUDA(cudaEventRecord(b->ev1, b->stream));
  VL_CUDA(cudaEventSynchronize(b->ev1));
  float ms = 0.f;
  VL_CUDA(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
  b->last_ms = ms;
  if (n_launches_out) *n_launches_out = launches;
  return 0;
}

int vl_batch_profile_step(vl_batch* b, int32_t n_updates, double* ms_out, int64_t* n_out) {
  if (!b) return fail("batch is NULL");
  if (!b->uploaded) return fail("vl_batch_profile_step before vl_batch_upload");
  if (!ms_out || !n_out || n_updates < 0) return fail("bad argument");
  VL_CUDA(cudaSetDevice(b->device));
  for (int i = 0; i < 4; ++i) { ms_out[i] = 0.0; n_out[i] = 0; }
  cudaEvent_t ev[2];
  for (auto& e : ev) VL_CUDA(cudaEventCreate(&e));
  auto timed = [&](int slot, auto&& fn) -> int {
    int64_t n = 0;
    VL_CUDA(cudaEventRecord(ev[0], b->stream));
    if (fn(&n)) return 1;
    VL_CUDA(cudaEventRecord(ev[1], b->stream));
    VL_CUDA(cudaEventSynchronize(ev[1]));
    float m = 0.f;
    VL_CUDA(cudaEventElapsedTime(&m, ev[0], ev[1]));
    if (n > 0) { ms_out[slot] += m; n_out[slot] += n; }
    return 0;
  };
  if (b->sched_dirty && schedule(b)) return 1;
  VL_CUDA(cudaStreamSynchronize(b->stream));
  for (int it = 0; it < n_updates; ++it)
    for (int phase = 0; phase < 2; ++phase)
      if (timed(phase, [&](int64_t* n) { return launch_phase(b, phase, n); })) return 1;
  if (sync_status(b, nullptr)) return 1;
  for (auto& e : ev) cudaEventDestroy(e);
  return 0;
}

int vl_batch_problem_stats(vl_batch* b, int32_t* out, int32_t n_problems) {
  if (!b || !out) return fail("NULL argument");
  if (n_problems != b->lay.n_problems) return fail("n_problems differs from the batch");
  for (int p = 0; p < n_problems; ++p) {
    const int4 st = b->pstat[p];
    out[4 * p + 0] = st.x; out[4 * p + 1] = st.y; out[4 * p + 2] = st.z; out[4 * p + 3] = st.w;
  }
  return 0;
}

static void fill_scalars(const VlState& S, double* out) {
  for (int i = 0; i < VL_SCALAR_STRIDE; ++i) out[i] = 0.0;
  out[VL_S_ELBO] = S.elbo; out[VL_S_GAMMA_A] = S.a; out[VL_S_GAMMA_B] = S.b;
  out[VL_S_MLG0] = S.mlg0; out[VL_S_MLG1] = S.mlg1;
  out[VL_S_THETA_C] = S.c; out[VL_S_THETA_D] = S.d; out[VL_S_MLT0] = S.mlt0; out[VL_S_MLT1] = S.mlt1;
  out[VL_S_DSUM_ALPHA] = S.dsum_alpha; out[VL_S_DSUM_AB] = S.dsum_ab; out[VL_S_DSUM_CD] = S.dsum_cd;
  out[VL_S_TERM_DATA] = S.t_data; out[VL_S_TERM_PI] = S.t_pi;
  out[VL_S_TERM_CLUSTER] = S.t_cluster; out[VL_S_TERM_HAPLO] = S.t_haplo;
}

int vl_batch_fetch(vl_batch* b, vl_window_result* results, int32_t n_windows) {
  if (!b) return fail("batch is NULL");
  if (!results) return fail("results is NULL");
  const Layout& L = b->lay;
  if (n_windows != (int)L.wins.size()) return fail("n_windows differs from the batch");
  VL_CUDA(cudaSetDevice(b->device));
  cudaStream_t s = b->stream;
  static const bool trace = std::getenv("VL_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (trace) fprintf(stderr, "fetch: %-28s at %.2f ms\n", what,
                       std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
  };
  VL_CUDA(cudaMemcpyAsync(b->states.data(), b->d_states(), b->states.size() * sizeof(VlState), cudaMemcpyDeviceToHost, s));
  VL_CUDA(cudaStreamSynchronize(s));
  lap("states on the host");
  for (int i = 0; i < n_windows; ++i) {
    const WindowLayout& W = L.wins[i];
    vl_window_result& R = results[i];
    int best = 0, max_hist = 0;
    for (int r = 0; r < W.R; ++r) {
      const VlState& S = b->states[W.first_problem + r];
      if (R.elbo) R.elbo[r] = S.elbo;
      if (R.n_iterations) R.n_iterations[r] = S.iter;
      if (R.n_updates) R.n_updates[r] = S.n_updates;
      if (R.status) R.status[r] = S.status;
      if (R.converged) R.converged[r] = S.converged;
      const int hl = std::min(S.n_hist, W.max_iter);
      if (R.history_len) R.history_len[r] = hl;
      max_hist = std::max(max_hist, hl);
      // stable descending sort of run_dpm_mfa.py:83-95: first maximum wins; NaN never wins
      const double eb = b->states[W.first_problem + best].elbo;
      if (r > 0 && (S.elbo > eb || (std::isnan(eb) && !std::isnan(S.elbo)))) best = r;
    }
    R.best_restart = best;
    if (R.history && max_hist > 0) {
      if (R.history_stride < max_hist) return fail("history_stride smaller than the longest ELBO history");
      VL_CUDA(cudaMemcpy2DAsync(R.history, (size_t)R.history_stride * 8, b->ws + W.rs[0].hist, (size_t)W.max_iter * 8,
                                (size_t)max_hist * 8, W.R, cudaMemcpyDeviceToHost, s));
    }
    const int sr = (R.state_restart < 0) ? best : R.state_restart;
    if (sr >= W.R) return fail("state_restart out of range");
    const RestartLayout& RL = W.rs[sr];
    const int p = W.first_problem + sr;
    const VlState& S = b->states[p];
    if (R.mean_haplo) {
      const size_t total = (size_t)W.K * W.L * VL_N_BASES;
      VL_CUDA(vl_launch_export_haplo(b->d_probs() + p, b->dev<double>(L.export_h), total, s));
      VL_CUDA(cudaMemcpyAsync(R.mean_haplo, b->ws + L.export_h, total * 8, cudaMemcpyDeviceToHost, s));
    }
    if (R.mean_cluster) {
      const size_t total = (size_t)W.N * W.K;
      VL_CUDA(vl_launch_export_cluster(b->d_probs() + p, b->dev<double>(L.export_z), total, s));
      VL_CUDA(cudaMemcpyAsync(R.mean_cluster, b->ws + L.export_z, total * 8, cudaMemcpyDeviceToHost, s));
    }
    if (R.alpha) VL_CUDA(cudaMemcpyAsync(R.alpha, b->ws + RL.alpha, (size_t)W.K * 8, cudaMemcpyDeviceToHost, s));
    if (R.mean_log_pi) VL_CUDA(cudaMemcpyAsync(R.mean_log_pi, b->ws + RL.mlp, (size_t)W.K * 8, cudaMemcpyDeviceToHost, s));
    if (R.scalars) fill_scalars(S, R.scalars);
    // the export buffers are reused by the next window: the next export kernel is ordered after
    // these copies on the same stream
  }
  lap("exports and copies enqueued");
  VL_CUDA(cudaStreamSynchronize(s));
  lap("copies done");
  return 0;
}

int vl_batch_set_state(vl_batch* b, int32_t window, int32_t restart, const double* mean_cluster,
                       const double* mean_log_pi, const double* scalars, int32_t iter) {
  if (!b) return fail("batch is NULL");
  if (!b->uploaded) return fail("vl_batch_set_state before vl_batch_upload");
  const Layout& L = b->lay;
  if (window < 0 || window >= (int)L.wins.size()) return fail("window out of range");
  const WindowLayout& W = L.wins[window];
  if (restart < 0 || restart >= W.R) return fail("restart out of range");
  if (!mean_cluster || !mean_log_pi || !scalars) return fail("NULL state pointer");
  VL_CUDA(cudaSetDevice(b->device));
  cudaStream_t s = b->stream;
  const RestartLayout& RL = W.rs[restart];
  const int p = W.first_problem + restart;
  // the state arrives in cluster order: back to the identity layout
  std::vector<int> lay, gat;
  identity_layout(W.K, W.K8, lay, gat);
  std::vector<double> lp(W.K8, 0.0);
  std::memcpy(lp.data(), mean_log_pi, (size_t)W.K * 8);
  VL_CUDA(cudaMemsetAsync(b->ws + RL.z, 0, (size_t)W.Npad * W.KS * 8, s));
  VL_CUDA(cudaMemcpy2DAsync(b->ws + RL.z, (size_t)W.KS * 8, mean_cluster, (size_t)W.K * 8, (size_t)W.K * 8, W.N, cudaMemcpyHostToDevice, s));
  VL_CUDA(vl_launch_scale_rows(b->dev<double>(RL.z), b->dev<double>(W.w), b->dev<double>(RL.wz), W.Npad, W.KS, s));
  VL_CUDA(cudaMemcpyAsync(b->ws + RL.mlp, lp.data(), (size_t)W.K8 * 8, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemcpyAsync(b->ws + RL.mlp_lay, lp.data(), (size_t)W.K8 * 8, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemcpyAsync(b->ws + RL.lay, lay.data(), (size_t)W.K8 * 4, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemcpyAsync(b->ws + RL.lay_z, lay.data(), (size_t)W.K8 * 4, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemcpyAsync(b->ws + RL.gat, gat.data(), (size_t)W.K8 * 4, cudaMemcpyHostToDevice, s));
  VlState S;
  VL_CUDA(cudaMemcpyAsync(&S, b->d_states() + p, sizeof S, cudaMemcpyDeviceToHost, s));
  VL_CUDA(cudaStreamSynchronize(s));
  S.mlg0 = scalars[VL_S_MLG0]; S.mlg1 = scalars[VL_S_MLG1];
  S.mlt0 = scalars[VL_S_MLT0]; S.mlt1 = scalars[VL_S_MLT1];
  S.dsum_alpha = scalars[VL_S_DSUM_ALPHA]; S.dsum_ab = scalars[VL_S_DSUM_AB]; S.dsum_cd = scalars[VL_S_DSUM_CD];
  S.iter = iter; S.converged = 0; S.status = VL_STATUS_RUNNING; S.active = 1;
  S.n_hist = 0; S.n_updates = 0; S.h_last = 0.0; S.h_prev = 0.0;
  S.ktl = W.KT; S.ktl_z = W.KT; S.nlay = W.K; S.ghost = -1; S.nlay_z = W.K; S.ghost_z = -1;
  S.ghost_mult = 0.0; S.stalled = 0; S.skip = b->skip_dead;
  S.ticket = 0; S.pub = 0; S.ktring[0] = W.KT; S.ktring[1] = W.KT;
  VL_CUDA(cudaMemcpyAsync(b->d_states() + p, &S, sizeof S, cudaMemcpyHostToDevice, s));
  const int4 st = make_int4(1, W.KT, W.K, 0);
  VL_CUDA(cudaMemcpyAsync(b->d_pstat() + p, &st, sizeof st, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaStreamSynchronize(s));
  b->pstat[p] = st;
  b->sched_dirty = true;
  return 0;
}

int vl_batch_merge_haplo(vl_batch* b, int32_t window, int32_t restart, const int32_t* group_of,
                         int32_t n_groups, double* merged_cluster_out, double* merged_haplo_out) {
  if (!b) return fail("batch is NULL");
  if (!b->uploaded) return fail("vl_batch_merge_haplo before vl_batch_upload");
  const Layout& L = b->lay;
  if (window < 0 || window >= (int)L.wins.size()) return fail("window out of range");
  const WindowLayout& W = L.wins[window];
  if (restart < 0 || restart >= W.R) return fail("restart out of range");
  if (!group_of || n_groups <= 0 || n_groups > W.K) return fail("bad cluster grouping");
  for (int k = 0; k < W.K; ++k)
    if (group_of[k] < 0 || group_of[k] >= n_groups) return fail("group_of entry out of range");
  VL_CUDA(cudaSetDevice(b->device));
  cudaStream_t s = b->stream;
  const int p = W.first_problem + restart;
  const int G = n_groups, KTg = (G + 7) / 8, KSg = vl_ks(KTg);
  VL_CUDA(cudaMemcpyAsync(b->ws + L.tmp_group, group_of, (size_t)W.K * sizeof(int), cudaMemcpyHostToDevice, s));
  VL_CUDA(vl_launch_merge(b->d_probs() + p, b->dev<double>(L.tmp_z), b->dev<double>(L.tmp_wz), b->dev<int>(L.tmp_group),
                          G, KSg, (size_t)W.Npad * KSg, s));
  // a one-problem batch: same window tensors, merged responsibilities in an identity layout
  // over the G groups, scratch outputs
  VlProblem T = b->probs[p];
  T.K = G; T.KT = KTg;
  T.n_hsplit = 1; T.tsplit = nullptr; T.hsplit_count = nullptr;      // one CTA per column tile here
  T.z = b->dev<double>(L.tmp_z); T.wz = b->dev<double>(L.tmp_wz);
  T.h = b->dev<double>(L.tmp_h); T.g = b->dev<double>(L.tmp_g); T.gm = b->dev<double>(L.tmp_gm);
  T.hpart = b->dev<double>(L.tmp_hpart);
  T.lay = b->dev<int>(L.tmp_lay); T.lay_z = b->dev<int>(L.tmp_lay); T.gat = b->dev<int>(L.tmp_gat);
  T.mlp_lay = b->dev<double>(L.tmp_mlp);
  T.st = b->dev<VlState>(L.tmp_state);
  VlState S;
  VL_CUDA(cudaMemcpyAsync(&S, b->d_states() + p, sizeof S, cudaMemcpyDeviceToHost, s));
  VL_CUDA(cudaStreamSynchronize(s));
  S.active = 1;
  S.ktl = KTg; S.ktl_z = KTg; S.nlay = G; S.ghost = -1; S.nlay_z = G; S.ghost_z = -1; S.ghost_mult = 0.0;
  S.ktring[0] = KTg; S.ktring[1] = KTg;
  std::vector<int> lay, gat;
  identity_layout(G, 8 * KTg, lay, gat);
  std::vector<int2> tiles(W.n_htiles);
  for (int t = 0; t < W.n_htiles; ++t) tiles[t] = make_int2(0, t);
  VL_CUDA(cudaMemcpyAsync(b->ws + L.tmp_lay, lay.data(), lay.size() * 4, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemcpyAsync(b->ws + L.tmp_gat, gat.data(), gat.size() * 4, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemcpyAsync(b->ws + L.tmp_state, &S, sizeof S, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemcpyAsync(b->ws + L.tmp_prob, &T, sizeof T, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemcpyAsync(b->ws + L.tmp_tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, s));
  VL_CUDA(vl_launch_haplo(vl_class_of_tiles(KTg), W.mode, 0, b->dev<VlProblem>(L.tmp_prob), b->dev<int2>(L.tmp_tiles),
                          W.n_htiles, s));
  if (merged_haplo_out) {
    const size_t total = (size_t)G * W.L * VL_N_BASES;
    VL_CUDA(vl_launch_export_haplo(b->dev<VlProblem>(L.tmp_prob), b->dev<double>(L.export_h), total, s));
    VL_CUDA(cudaMemcpyAsync(merged_haplo_out, b->ws + L.export_h, total * 8, cudaMemcpyDeviceToHost, s));
  }
  if (merged_cluster_out)
    VL_CUDA(cudaMemcpy2DAsync(merged_cluster_out, (size_t)G * 8, b->ws + L.tmp_z, (size_t)KSg * 8,
                              (size_t)G * 8, W.N, cudaMemcpyDeviceToHost, s));
  VL_CUDA(cudaStreamSynchronize(s));
  return 0;
}

int vl_debug_occupancy(int32_t device, int32_t kernel_class, int32_t mode, int32_t deep, int32_t* ctas_per_sm,
                       int32_t* smem_bytes, int32_t* registers) {
  if (!ctas_per_sm || !smem_bytes || !registers) return fail("NULL output");
  if (kernel_class < 0 || kernel_class >= VL_N_CLASSES) return fail("kernel_class out of range");
  VL_CUDA(cudaSetDevice(device));
  int c = 0, sm = 0, r = 0;
  VL_CUDA(vl_sched_occupancy(kernel_class, mode, 0, deep, 0, &c, &sm, &r));
  *ctas_per_sm = c; *smem_bytes = sm; *registers = r;
  return 0;
}

int vl_debug_exp(int32_t device, const double* x, double* out, int32_t n) {
  if (!x || !out || n <= 0) return fail("bad argument");
  VL_CUDA(cudaSetDevice(device));
  double* d = nullptr;
  VL_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), (size_t)n * 16));
  cudaError_t e = cudaMemcpy(d, x, (size_t)n * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = vl_launch_exp_test(d, d + n, n, nullptr);
  if (e == cudaSuccess) e = cudaMemcpy(out, d + n, (size_t)n * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  VL_CUDA(e);
  return 0;
}

int vl_debug_fp64_peak(int32_t device, double* tflops) {
  if (!tflops) return fail("bad argument");
  VL_CUDA(cudaSetDevice(device));
  int sms = 0;
  VL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int blocks = 2 * sms, iters = 4096;
  double* d = nullptr;
  VL_CUDA(cudaMalloc(reinterpret_cast<void**>(&d), (size_t)blocks * 256 * sizeof(double)));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 5 && e == cudaSuccess; ++r) {        // (the first repetitions warm the clocks up)
    e = cudaEventRecord(e0, nullptr);
    if (e == cudaSuccess) e = vl_launch_dmma_peak(d, blocks, iters, nullptr);
    if (e == cudaSuccess) e = cudaEventRecord(e1, nullptr);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e == cudaSuccess && r >= 2 && ms < best) best = ms;
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(d);
  VL_CUDA(e);
  const double flops = 2.0 * 8 * 8 * 4 * 8.0 * iters * (256 / 32) * (double)blocks;
  *tflops = flops / ((double)best * 1e-3) * 1e-12;
  return 0;
}

int vl_debug_counters(uint64_t* out36, int32_t reset) {
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "counter width");
  for (int i = 0; i < 16; ++i) out36[i] = 0;
  int rc = vl_debug_general_cycles(reinterpret_cast<unsigned long long*>(out36 + 16), reset);
  if (rc == 2) return fail("library built without -DVL_RS_PROFILE");
  return rc ? fail("cudaMemcpyFromSymbol failed") : 0;
}

double vl_batch_last_device_ms(const vl_batch* b) { return b ? b->last_ms : 0.0; }

#pragma GCC visibility pop
}  // extern "C"
