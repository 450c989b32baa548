// Host-only part of the library: a window's read file -> compact tensors.
//
// Restates load_and_process_reads of the reference (use_quality_scores/preparation.py:24-90,
// learn_error_params/preparation.py:51-94) as native code: FASTA records, identical read strings
// collapsed into the row of their first occurrence (weight = multiplicity), qualities averaged with the
// reference's running mean in file order, Q == 0 -> 2, bases -> codes.  SURVEY.md 8(f) row 1: for a
// 5 000-read window the per-record Python loops cost more host time than the whole CAVI run costs
// the GPU.  No CUDA call in this file.
#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "../../include/viloca_b200.h"
#include "vl_launch.h"

struct vl_reads {
  const char* text = nullptr;            // the caller's buffer (must outlive the handle)
  size_t n_bytes = 0;
  int length = 0;
  bool unique = true;
  std::vector<int64_t> id_off, id_len;   // per raw read: first token of the header
  std::vector<int64_t> seq_off;          // per raw read: its sequence (one line, `length` characters)
  std::vector<int32_t> row_of;           // per raw read: its row
  std::vector<int32_t> first_read;       // per row: the raw read that opened it
};

namespace {

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int vl_reads_parse(const char* fasta, size_t n_bytes, int32_t unique, vl_reads** out, int32_t* n_raw_out,
                   int32_t* n_rows_out, int32_t* length_out) {
  if (!fasta || !out || !n_raw_out || !n_rows_out || !length_out) return vl_set_error("vl_reads_parse: NULL argument");
  auto* R = new vl_reads;
  R->text = fasta; R->n_bytes = n_bytes; R->unique = unique != 0;
  size_t p = 0;
  bool have_title = false, have_seq = false;
  auto fail = [&](const char* msg) { delete R; return vl_set_error(msg); };
  while (p < n_bytes) {
    size_t e = p;
    while (e < n_bytes && fasta[e] != '\n') ++e;
    size_t le = e;                                         // line = [p, le) without \r\n
    while (le > p && (fasta[le - 1] == '\r')) --le;
    if (le > p && fasta[p] == '>') {
      if (have_title && !have_seq) return fail("record without a sequence line: use the generic reader");
      size_t a = p + 1;
      while (a < le && is_space(fasta[a])) ++a;            // str.split(None, 1)[0]
      size_t b = a;
      while (b < le && !is_space(fasta[b])) ++b;
      R->id_off.push_back((int64_t)a); R->id_len.push_back((int64_t)(b - a));
      have_title = true; have_seq = false;
    } else if (have_title) {
      size_t a = p, b = le;                                // str.strip()
      while (a < b && is_space(fasta[a])) ++a;
      while (b > a && is_space(fasta[b - 1])) --b;
      if (have_seq) {
        if (b > a) return fail("multi-line sequences: use the generic reader");
      } else {
        const int len = (int)(b - a);
        if (R->seq_off.empty()) R->length = len;
        else if (len != R->length) return fail("reads of one window must have equal length");
        R->seq_off.push_back((int64_t)a);
        have_seq = true;
      }
    }
    p = e + 1;
  }
  if (have_title && !have_seq) return fail("record without a sequence line: use the generic reader");
  const int n_raw = (int)R->id_off.size();
  if (n_raw == 0) return fail("window without reads");
  R->row_of.resize(n_raw);
  if (R->unique) {
    std::unordered_map<std::string_view, int32_t> rows;
    rows.reserve((size_t)n_raw * 2);
    for (int i = 0; i < n_raw; ++i) {
      const std::string_view key(fasta + R->seq_off[i], (size_t)R->length);
      auto it = rows.find(key);
      if (it == rows.end()) {
        const int32_t r = (int32_t)R->first_read.size();
        rows.emplace(key, r);
        R->first_read.push_back(i);
        R->row_of[i] = r;
      } else {
        R->row_of[i] = it->second;
      }
    }
  } else {
    R->first_read.resize(n_raw);
    for (int i = 0; i < n_raw; ++i) { R->row_of[i] = i; R->first_read[i] = i; }
  }
  *out = R;
  *n_raw_out = n_raw;
  *n_rows_out = (int32_t)R->first_read.size();
  *length_out = R->length;
  return 0;
}

int vl_reads_fill(const vl_reads* R, const char* alphabet, const void* qualities, int32_t quality_kind,
                  uint8_t* codes, double* weights, double* mean_qualities, int32_t* row_of, int64_t* id_spans,
                  int64_t* seq_spans) {
  if (!R || !alphabet || !codes || !weights) return vl_set_error("vl_reads_fill: NULL argument");
  if (quality_kind != 0 && (!qualities || !mean_qualities)) return vl_set_error("vl_reads_fill: qualities without output");
  const int n_raw = (int)R->row_of.size(), n_rows = (int)R->first_read.size(), L = R->length;
  uint8_t lut[256];
  std::memset(lut, 255, sizeof lut);
  for (int i = 0; alphabet[i] != 0 && i < 250; ++i) lut[(unsigned char)alphabet[i]] = (uint8_t)i;
  for (int r = 0; r < n_rows; ++r) {
    const char* s = R->text + R->seq_off[R->first_read[r]];
    uint8_t* c = codes + (size_t)r * L;
    for (int l = 0; l < L; ++l) c[l] = lut[(unsigned char)s[l]];
    weights[r] = 0.0;
    if (seq_spans) { seq_spans[2 * r] = R->seq_off[R->first_read[r]]; seq_spans[2 * r + 1] = L; }
  }
  const int64_t* qi = quality_kind == 1 ? static_cast<const int64_t*>(qualities) : nullptr;
  const double* qd = quality_kind == 2 ? static_cast<const double*>(qualities) : nullptr;
  for (int i = 0; i < n_raw; ++i) {
    const int r = R->row_of[i];
    if (quality_kind != 0) {
      double* q = mean_qualities + (size_t)r * L;
      const double w = weights[r];
      if (w == 0.0) {
        for (int l = 0; l < L; ++l) q[l] = qi ? (double)qi[(size_t)i * L + l] : qd[(size_t)i * L + l];
      } else {
        // the reference's running mean, in file order: (q * w + q_new) / (w + 1)   (preparation.py:71-74)
        const double w1 = w + 1.0;
        for (int l = 0; l < L; ++l) {
          const double qn = qi ? (double)qi[(size_t)i * L + l] : qd[(size_t)i * L + l];
          q[l] = (q[l] * w + qn) / w1;
        }
      }
    }
    weights[r] += 1.0;
    if (row_of) row_of[i] = r;
    if (id_spans) { id_spans[2 * i] = R->id_off[i]; id_spans[2 * i + 1] = R->id_len[i]; }
  }
  if (quality_kind != 0)
    for (size_t k = 0, n = (size_t)n_rows * L; k < n; ++k)
      if (mean_qualities[k] == 0.0) mean_qualities[k] = 2.0;           // preparation.py:88
  return 0;
}

void vl_reads_free(vl_reads* R) { delete R; }

}  // extern "C"
#pragma GCC visibility pop
