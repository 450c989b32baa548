// C ABI of the CAVI engine (include/viloca_b200.h): workspace layout, host<->device
// transfers, the iteration driver, and result export.  No process-global state other than a
// thread-local error string.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
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
  size_t z, h, zpart, hpart, alpha, mlp, hist;
};

struct WindowLayout {
  int mode, N, L, K, R, max_iter;
  int Npad, Ls, KT, KS, K8, n_ztiles, n_htiles;
  double alpha0, conv_thres;
  size_t codes, lt, lo, w, mcount, ref;
  std::vector<RestartLayout> rs;
  int first_problem;
};

struct Layout {
  std::vector<WindowLayout> wins;
  int n_problems = 0;
  size_t data_begin = 0, data_end = 0;   // zero-filled on upload
  size_t probs = 0, states = 0, done = 0;
  size_t htiles = 0, ztiles = 0;         // tile maps of all classes, back to back
  size_t export_buf = 0;                 // max K*L*B doubles
  size_t tmp_z = 0, tmp_h = 0, tmp_hpart = 0, tmp_group = 0, tmp_prob = 0, tmp_state = 0, tmp_tiles = 0;
  size_t total = 0;
  size_t n_htiles_total = 0, n_ztiles_total = 0;
};

struct KernelClass {
  int kt, mode;
  std::vector<int2> htiles, ztiles;
  size_t h_off = 0, z_off = 0;   // element offsets into the device tile maps
};

int build_layout(const vl_window_desc* w, int n, Layout& lay, std::string& err) {
  if (n <= 0 || !w) { err = "no windows"; return 1; }
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes); return o; };
  lay.wins.resize(n);
  lay.n_problems = 0;
  lay.data_begin = 0;
  size_t max_export = 0, max_z = 0, max_h = 0;
  int max_htiles = 0;
  for (int i = 0; i < n; ++i) {
    const vl_window_desc& d = w[i];
    WindowLayout& W = lay.wins[i];
    char buf[256];
    if (d.mode != VL_MODE_UQS && d.mode != VL_MODE_LEP) { snprintf(buf, sizeof buf, "window %d: bad mode %d", i, d.mode); err = buf; return 1; }
    if (d.n_reads <= 0 || d.length <= 0 || d.n_clusters <= 0 || d.n_starts <= 0 || d.max_iter <= 0) {
      snprintf(buf, sizeof buf, "window %d: n_reads, length, n_clusters, n_starts and max_iter must be positive", i);
      err = buf; return 1;
    }
    const int kt = vl_supported_kt(d.n_clusters);
    if (kt < 0) { snprintf(buf, sizeof buf, "window %d: n_clusters=%d exceeds VL_MAX_CLUSTERS=%d", i, d.n_clusters, VL_MAX_CLUSTERS); err = buf; return 1; }
    W.mode = d.mode; W.N = d.n_reads; W.L = d.length; W.K = d.n_clusters; W.R = d.n_starts;
    W.max_iter = d.max_iter; W.alpha0 = d.alpha0; W.conv_thres = d.conv_thres;
    W.Npad = round_up(W.N, VL_ZT_READS);
    W.Ls = round_up(W.L, VL_HT_POS);
    W.KT = kt; W.KS = vl_ks(kt); W.K8 = vl_k8(kt);
    W.n_ztiles = W.Npad / VL_ZT_READS;
    W.n_htiles = W.Ls / VL_HT_POS;
    const size_t nl = (size_t)W.Npad * W.Ls;
    W.codes = take(nl);
    if (W.mode == VL_MODE_UQS) { W.lt = take(nl * 8); W.lo = take(nl * 8); } else { W.lt = W.lo = 0; }
    W.w = take((size_t)W.Npad * 8);
    W.mcount = take((size_t)W.Npad * 8);
    W.ref = take((size_t)W.Ls);
    W.rs.resize(W.R);
    W.first_problem = lay.n_problems;
    for (int r = 0; r < W.R; ++r) {
      RestartLayout& RL = W.rs[r];
      RL.z = take((size_t)W.Npad * W.KS * 8);
      RL.h = take((size_t)W.Ls * VL_B * W.KS * 8);
      RL.zpart = take((size_t)W.n_ztiles * (W.K8 + VL_ZP_EXTRA) * 8);
      RL.hpart = take((size_t)W.n_htiles * VL_HP_STRIDE * 8);
      RL.alpha = take((size_t)W.K8 * 8);
      RL.mlp = take((size_t)W.K8 * 8);
    }
    // history_elbo of the R restarts, contiguous so that one 2-D copy fetches them
    const size_t hist0 = take((size_t)W.R * W.max_iter * 8);
    for (int r = 0; r < W.R; ++r) W.rs[r].hist = hist0 + (size_t)r * W.max_iter * 8;
    lay.n_problems += W.R;
    max_export = std::max(max_export, (size_t)W.K * W.L * VL_B * 8);
    max_z = std::max(max_z, (size_t)W.Npad * W.KS * 8);
    max_h = std::max(max_h, (size_t)W.Ls * VL_B * W.KS * 8);
    max_htiles = std::max(max_htiles, W.n_htiles);
    lay.n_htiles_total += (size_t)W.n_htiles * W.R;
    lay.n_ztiles_total += (size_t)W.n_ztiles * W.R;
  }
  lay.data_end = off;
  lay.probs = take((size_t)lay.n_problems * sizeof(VlProblem));
  lay.states = take((size_t)lay.n_problems * sizeof(VlState));
  lay.done = take(64);
  lay.htiles = take(lay.n_htiles_total * sizeof(int2));
  lay.ztiles = take(lay.n_ztiles_total * sizeof(int2));
  lay.export_buf = take(max_export);
  lay.tmp_z = take(max_z);
  lay.tmp_h = take(max_h);
  lay.tmp_hpart = take((size_t)max_htiles * VL_HP_STRIDE * 8);
  lay.tmp_group = take(VL_MAX_CLUSTERS * sizeof(int));
  lay.tmp_prob = take(sizeof(VlProblem));
  lay.tmp_state = take(sizeof(VlState));
  lay.tmp_tiles = take((size_t)max_htiles * sizeof(int2));
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
  std::vector<KernelClass> classes;
  std::vector<VlState> states;         // host mirror filled by fetch / set_state
  int* h_done = nullptr;               // pinned
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double last_ms = 0.0;
  bool uploaded = false;

  template <typename T> T* dev(size_t off) const { return reinterpret_cast<T*>(ws + off); }
  VlProblem* d_probs() const { return dev<VlProblem>(lay.probs); }
  VlState* d_states() const { return dev<VlState>(lay.states); }
  int* d_done() const { return dev<int>(lay.done); }
};

namespace {

int run_iterations(vl_batch* b, int n_iter, int64_t* launches) {
  for (int it = 0; it < n_iter; ++it) {
    for (const KernelClass& c : b->classes) {
      VL_CUDA(vl_launch_contractions(c.kt, c.mode, b->d_probs(), b->dev<int2>(b->lay.htiles) + c.h_off,
                                     (int)c.htiles.size(), b->dev<int2>(b->lay.ztiles) + c.z_off,
                                     (int)c.ztiles.size(), true, true, b->stream));
      *launches += 2;
    }
    VL_CUDA(vl_launch_scalar(b->d_probs(), b->lay.n_problems, b->d_done(), b->stream));
    *launches += 1;
  }
  return 0;
}

}  // namespace

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
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { delete b; return fail(std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }

  // problem table + tile maps, grouped by kernel class (cluster-tile count, mode); within a
  // class the costliest windows come first so that the long CTAs start early
  const Layout& L = b->lay;
  b->probs.resize(L.n_problems);
  b->states.resize(L.n_problems);
  std::map<std::pair<int, int>, int> class_of;
  std::vector<int> order(L.wins.size());
  for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
    const WindowLayout& A = L.wins[x]; const WindowLayout& B = L.wins[y];
    return (double)A.Npad * A.Ls > (double)B.Npad * B.Ls;
  });
  for (size_t i = 0; i < L.wins.size(); ++i) {
    const WindowLayout& W = L.wins[i];
    for (int r = 0; r < W.R; ++r) {
      VlProblem& P = b->probs[W.first_problem + r];
      P.mode = W.mode; P.N = W.N; P.L = W.L; P.K = W.K;
      P.Npad = W.Npad; P.Ls = W.Ls; P.KS = W.KS; P.KT = W.KT;
      P.n_ztiles = W.n_ztiles; P.n_htiles = W.n_htiles; P.window = (int)i; P.restart = r;
      P.codes = b->dev<uint8_t>(W.codes);
      P.lt = (W.mode == VL_MODE_UQS) ? b->dev<double>(W.lt) : nullptr;
      P.lo = (W.mode == VL_MODE_UQS) ? b->dev<double>(W.lo) : nullptr;
      P.w = b->dev<double>(W.w);
      P.mcount = b->dev<double>(W.mcount);
      P.ref = b->dev<uint8_t>(W.ref);
      const RestartLayout& RL = W.rs[r];
      P.z = b->dev<double>(RL.z); P.h = b->dev<double>(RL.h);
      P.zpart = b->dev<double>(RL.zpart); P.hpart = b->dev<double>(RL.hpart);
      P.alpha = b->dev<double>(RL.alpha); P.mlp = b->dev<double>(RL.mlp);
      P.hist = b->dev<double>(RL.hist);
      P.st = b->d_states() + (W.first_problem + r);
    }
  }
  for (int wi : order) {
    const WindowLayout& W = L.wins[wi];
    auto key = std::make_pair(W.KT, W.mode);
    if (!class_of.count(key)) { class_of[key] = (int)b->classes.size(); b->classes.push_back(KernelClass{W.KT, W.mode}); }
    KernelClass& C = b->classes[class_of[key]];
    for (int r = 0; r < W.R; ++r) {
      const int p = W.first_problem + r;
      for (int t = 0; t < W.n_htiles; ++t) C.htiles.push_back(make_int2(p, t));
      for (int t = 0; t < W.n_ztiles; ++t) C.ztiles.push_back(make_int2(p, t));
    }
  }
  size_t ho = 0, zo = 0;
  for (KernelClass& C : b->classes) { C.h_off = ho; C.z_off = zo; ho += C.htiles.size(); zo += C.ztiles.size(); }

#define VL_CREATE_CUDA(expr)                                                                    \
  do {                                                                                          \
    cudaError_t e2_ = (expr);                                                                   \
    if (e2_ != cudaSuccess) {                                                                   \
      std::string m_ = std::string("CUDA error in vl_batch_create: ") + cudaGetErrorString(e2_); \
      vl_batch_destroy(b);                                                                      \
      return fail(m_);                                                                          \
    }                                                                                           \
  } while (0)
  VL_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void**>(&b->h_done), 64));
  VL_CREATE_CUDA(cudaEventCreate(&b->ev0));
  VL_CREATE_CUDA(cudaEventCreate(&b->ev1));
  VL_CREATE_CUDA(cudaMemcpyAsync(b->d_probs(), b->probs.data(), b->probs.size() * sizeof(VlProblem),
                                 cudaMemcpyHostToDevice, b->stream));
  for (const KernelClass& C : b->classes) {
    VL_CREATE_CUDA(cudaMemcpyAsync(b->dev<int2>(L.htiles) + C.h_off, C.htiles.data(), C.htiles.size() * sizeof(int2),
                                   cudaMemcpyHostToDevice, b->stream));
    VL_CREATE_CUDA(cudaMemcpyAsync(b->dev<int2>(L.ztiles) + C.z_off, C.ztiles.data(), C.ztiles.size() * sizeof(int2),
                                   cudaMemcpyHostToDevice, b->stream));
  }
  VL_CREATE_CUDA(cudaStreamSynchronize(b->stream));
#undef VL_CREATE_CUDA
  *out = b;
  return 0;
}

void vl_batch_destroy(vl_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->h_done) cudaFreeHost(b->h_done);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  delete b;
}

int vl_batch_upload(vl_batch* b, const vl_window_desc* windows, int32_t n_windows) {
  if (!b) return fail("batch is NULL");
  const Layout& L = b->lay;
  if (n_windows != (int)L.wins.size()) return fail("n_windows differs from the batch");
  VL_CUDA(cudaSetDevice(b->device));
  cudaStream_t s = b->stream;
  VL_CUDA(cudaMemsetAsync(b->ws + L.data_begin, 0, L.data_end - L.data_begin, s));
  for (int i = 0; i < n_windows; ++i) {
    const vl_window_desc& d = windows[i];
    const WindowLayout& W = L.wins[i];
    if (d.mode != W.mode || d.n_reads != W.N || d.length != W.L || d.n_clusters != W.K || d.n_starts != W.R)
      return fail("window shape differs from the one the batch was created with");
    if (!d.codes || !d.weights || !d.ref_codes || !d.init_cluster || !d.init_scalars)
      return fail("window input pointer is NULL");
    if (W.mode == VL_MODE_UQS && (!d.log_hit || !d.log_off)) return fail("uqs window without log_hit/log_off");
    // padding of codes must read as 'N': fill, then copy the real rows over it
    VL_CUDA(cudaMemsetAsync(b->ws + W.codes, VL_CODE_N, (size_t)W.Npad * W.Ls, s));
    VL_CUDA(cudaMemsetAsync(b->ws + W.ref, VL_CODE_N, (size_t)W.Ls, s));
    VL_CUDA(cudaMemcpy2DAsync(b->ws + W.codes, W.Ls, d.codes, W.L, W.L, W.N, cudaMemcpyHostToDevice, s));
    if (W.mode == VL_MODE_UQS) {
      VL_CUDA(cudaMemcpy2DAsync(b->ws + W.lt, (size_t)W.Ls * 8, d.log_hit, (size_t)W.L * 8, (size_t)W.L * 8, W.N, cudaMemcpyHostToDevice, s));
      VL_CUDA(cudaMemcpy2DAsync(b->ws + W.lo, (size_t)W.Ls * 8, d.log_off, (size_t)W.L * 8, (size_t)W.L * 8, W.N, cudaMemcpyHostToDevice, s));
    }
    VL_CUDA(cudaMemcpyAsync(b->ws + W.w, d.weights, (size_t)W.N * 8, cudaMemcpyHostToDevice, s));
    VL_CUDA(cudaMemcpyAsync(b->ws + W.ref, d.ref_codes, (size_t)W.L, cudaMemcpyHostToDevice, s));
    VL_CUDA(vl_launch_mcount(b->dev<uint8_t>(W.codes), b->dev<double>(W.mcount), W.Npad, W.Ls, s));
    for (int r = 0; r < W.R; ++r) {
      VL_CUDA(cudaMemcpy2DAsync(b->ws + W.rs[r].z, (size_t)W.KS * 8, d.init_cluster + (size_t)r * W.N * W.K,
                                (size_t)W.K * 8, (size_t)W.K * 8, W.N, cudaMemcpyHostToDevice, s));
      VlState& S = b->states[W.first_problem + r];
      std::memset(&S, 0, sizeof S);
      const double* in = d.init_scalars + (size_t)r * VL_INIT_STRIDE;
      S.alpha0 = W.alpha0; S.thr = W.conv_thres;
      S.a0 = in[0]; S.b0 = in[1]; S.mlg0 = in[2]; S.mlg1 = in[3];
      S.c0 = in[4]; S.d0 = in[5]; S.mlt0 = in[6]; S.mlt1 = in[7];
      S.max_iter = W.max_iter;
    }
  }
  VL_CUDA(cudaMemcpyAsync(b->d_states(), b->states.data(), b->states.size() * sizeof(VlState), cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemsetAsync(b->d_done(), 0, 64, s));
  VL_CUDA(vl_launch_init(b->d_probs(), L.n_problems, s));
  VL_CUDA(cudaStreamSynchronize(s));
  b->uploaded = true;
  return 0;
}

int vl_batch_run(vl_batch* b, int64_t* n_launches_out) {
  if (!b) return fail("batch is NULL");
  if (!b->uploaded) return fail("vl_batch_run before vl_batch_upload");
  VL_CUDA(cudaSetDevice(b->device));
  int64_t launches = 0;
  const int chunk = 8;
  int max_iter = 0;
  for (const WindowLayout& W : b->lay.wins) max_iter = std::max(max_iter, W.max_iter);
  VL_CUDA(cudaEventRecord(b->ev0, b->stream));
  for (int done_iters = 0; done_iters < max_iter + chunk; done_iters += chunk) {
    if (run_iterations(b, chunk, &launches)) return 1;
    VL_CUDA(cudaMemcpyAsync(b->h_done, b->d_done(), sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    VL_CUDA(cudaStreamSynchronize(b->stream));
    if (*b->h_done >= b->lay.n_problems) break;
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
  if (run_iterations(b, n_updates, &launches)) return 1;
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
  for (int i = 0; i < 3; ++i) { ms_out[i] = 0.0; n_out[i] = 0; }
  cudaEvent_t ev[4];
  for (auto& e : ev) VL_CUDA(cudaEventCreate(&e));
  const Layout& L = b->lay;
  for (int it = 0; it < n_updates; ++it) {
    for (const KernelClass& c : b->classes) {
      VL_CUDA(cudaEventRecord(ev[0], b->stream));
      VL_CUDA(vl_launch_contractions(c.kt, c.mode, b->d_probs(), b->dev<int2>(L.htiles) + c.h_off, (int)c.htiles.size(),
                                     nullptr, 0, true, false, b->stream));
      VL_CUDA(cudaEventRecord(ev[1], b->stream));
      VL_CUDA(vl_launch_contractions(c.kt, c.mode, b->d_probs(), nullptr, 0, b->dev<int2>(L.ztiles) + c.z_off,
                                     (int)c.ztiles.size(), false, true, b->stream));
      VL_CUDA(cudaEventRecord(ev[2], b->stream));
      VL_CUDA(cudaEventSynchronize(ev[2]));
      float m0 = 0.f, m1 = 0.f;
      VL_CUDA(cudaEventElapsedTime(&m0, ev[0], ev[1]));
      VL_CUDA(cudaEventElapsedTime(&m1, ev[1], ev[2]));
      ms_out[0] += m0; ms_out[1] += m1; n_out[0] += 1; n_out[1] += 1;
    }
    VL_CUDA(cudaEventRecord(ev[2], b->stream));
    VL_CUDA(vl_launch_scalar(b->d_probs(), L.n_problems, b->d_done(), b->stream));
    VL_CUDA(cudaEventRecord(ev[3], b->stream));
    VL_CUDA(cudaEventSynchronize(ev[3]));
    float m2 = 0.f;
    VL_CUDA(cudaEventElapsedTime(&m2, ev[2], ev[3]));
    ms_out[2] += m2; n_out[2] += 1;
  }
  for (auto& e : ev) cudaEventDestroy(e);
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
  VL_CUDA(cudaMemcpyAsync(b->states.data(), b->d_states(), b->states.size() * sizeof(VlState), cudaMemcpyDeviceToHost, s));
  VL_CUDA(cudaStreamSynchronize(s));
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
    const VlState& S = b->states[W.first_problem + sr];
    if (R.mean_haplo) {
      VL_CUDA(vl_launch_export_haplo(b->dev<double>(RL.h), b->dev<double>(L.export_buf), W.K, W.L, W.KS, s));
      VL_CUDA(cudaMemcpyAsync(R.mean_haplo, b->ws + L.export_buf, (size_t)W.K * W.L * VL_B * 8, cudaMemcpyDeviceToHost, s));
      // export_buf is reused by the next window
      VL_CUDA(cudaStreamSynchronize(s));
    }
    if (R.mean_cluster)
      VL_CUDA(cudaMemcpy2DAsync(R.mean_cluster, (size_t)W.K * 8, b->ws + RL.z, (size_t)W.KS * 8, (size_t)W.K * 8, W.N, cudaMemcpyDeviceToHost, s));
    if (R.alpha) VL_CUDA(cudaMemcpyAsync(R.alpha, b->ws + RL.alpha, (size_t)W.K * 8, cudaMemcpyDeviceToHost, s));
    if (R.mean_log_pi) VL_CUDA(cudaMemcpyAsync(R.mean_log_pi, b->ws + RL.mlp, (size_t)W.K * 8, cudaMemcpyDeviceToHost, s));
    if (R.scalars) fill_scalars(S, R.scalars);
  }
  VL_CUDA(cudaStreamSynchronize(s));
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
  VL_CUDA(cudaMemcpy2DAsync(b->ws + RL.z, (size_t)W.KS * 8, mean_cluster, (size_t)W.K * 8, (size_t)W.K * 8, W.N, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemcpyAsync(b->ws + RL.mlp, mean_log_pi, (size_t)W.K * 8, cudaMemcpyHostToDevice, s));
  VlState S;
  VL_CUDA(cudaMemcpyAsync(&S, b->d_states() + p, sizeof S, cudaMemcpyDeviceToHost, s));
  VL_CUDA(cudaStreamSynchronize(s));
  if (S.active == 0 && S.status != VL_STATUS_RUNNING) {
    // revive a finished restart; the done counter must stay consistent
    VL_CUDA(cudaMemcpyAsync(b->h_done, b->d_done(), sizeof(int), cudaMemcpyDeviceToHost, s));
    VL_CUDA(cudaStreamSynchronize(s));
    *b->h_done -= 1;
    VL_CUDA(cudaMemcpyAsync(b->d_done(), b->h_done, sizeof(int), cudaMemcpyHostToDevice, s));
  }
  S.mlg0 = scalars[VL_S_MLG0]; S.mlg1 = scalars[VL_S_MLG1];
  S.mlt0 = scalars[VL_S_MLT0]; S.mlt1 = scalars[VL_S_MLT1];
  S.dsum_alpha = scalars[VL_S_DSUM_ALPHA]; S.dsum_ab = scalars[VL_S_DSUM_AB]; S.dsum_cd = scalars[VL_S_DSUM_CD];
  S.iter = iter; S.converged = 0; S.status = VL_STATUS_RUNNING; S.active = 1;
  S.n_hist = 0; S.n_updates = 0; S.h_last = 0.0; S.h_prev = 0.0;
  VL_CUDA(cudaMemcpyAsync(b->d_states() + p, &S, sizeof S, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaStreamSynchronize(s));
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
  VL_CUDA(cudaMemcpyAsync(b->ws + L.tmp_group, group_of, (size_t)W.K * sizeof(int), cudaMemcpyHostToDevice, s));
  VL_CUDA(vl_launch_merge(b->probs[p].z, b->dev<double>(L.tmp_z), b->dev<int>(L.tmp_group), W.Npad, W.K, n_groups, W.KS, s));
  // a one-problem batch: same window tensors, merged responsibilities, scratch outputs
  VlProblem T = b->probs[p];
  T.K = n_groups;
  T.z = b->dev<double>(L.tmp_z);
  T.h = b->dev<double>(L.tmp_h);
  T.hpart = b->dev<double>(L.tmp_hpart);
  T.st = b->dev<VlState>(L.tmp_state);
  VlState S;
  VL_CUDA(cudaMemcpyAsync(&S, b->d_states() + p, sizeof S, cudaMemcpyDeviceToHost, s));
  VL_CUDA(cudaStreamSynchronize(s));
  S.active = 1;
  std::vector<int2> tiles(W.n_htiles);
  for (int t = 0; t < W.n_htiles; ++t) tiles[t] = make_int2(0, t);
  VL_CUDA(cudaMemcpyAsync(b->ws + L.tmp_state, &S, sizeof S, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemcpyAsync(b->ws + L.tmp_prob, &T, sizeof T, cudaMemcpyHostToDevice, s));
  VL_CUDA(cudaMemcpyAsync(b->ws + L.tmp_tiles, tiles.data(), tiles.size() * sizeof(int2), cudaMemcpyHostToDevice, s));
  VL_CUDA(vl_launch_contractions(W.KT, W.mode, b->dev<VlProblem>(L.tmp_prob), b->dev<int2>(L.tmp_tiles), W.n_htiles,
                                 nullptr, 0, true, false, s));
  if (merged_haplo_out) {
    VL_CUDA(vl_launch_export_haplo(b->dev<double>(L.tmp_h), b->dev<double>(L.export_buf), n_groups, W.L, W.KS, s));
    VL_CUDA(cudaMemcpyAsync(merged_haplo_out, b->ws + L.export_buf, (size_t)n_groups * W.L * VL_B * 8, cudaMemcpyDeviceToHost, s));
  }
  if (merged_cluster_out)
    VL_CUDA(cudaMemcpy2DAsync(merged_cluster_out, (size_t)n_groups * 8, b->ws + L.tmp_z, (size_t)W.KS * 8,
                              (size_t)n_groups * 8, W.N, cudaMemcpyDeviceToHost, s));
  VL_CUDA(cudaStreamSynchronize(s));
  return 0;
}

double vl_batch_last_device_ms(const vl_batch* b) { return b ? b->last_ms : 0.0; }

#pragma GCC visibility pop
}  // extern "C"
