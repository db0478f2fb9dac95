// Host-side writer of printAlignment's two files:
//   <prefix>.txt   '%d ' per phone + blank line                         (image_phone_hmm_word_discoverer.py:643-645)
//   <prefix>.json  json.dump(aligns, f, indent=4, sort_keys=True)        (:647-648)
// byte-for-byte what CPython's json module emits for the reference's list of dicts, straight from
// the flat arrays the batched decode kernels return -- no per-pair Python objects, no float boxing.
// With EM at millions of pairs per second the reference's per-pair dict building + json.dump
// (~250 floats per pair, each on its own indented line) is what dominates an alignment dump.
//
// Floats are printed like float.__repr__: shortest round-trip digits (std::to_chars), exponent form
// iff decpt <= -4 or decpt > 16, at least two exponent digits, 'NaN' / 'Infinity' / '-Infinity'.
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "mwd_common.cuh"

namespace {

// float.__repr__ of v appended to out
inline void put_float(std::string& out, double v) {
  if (std::isnan(v)) { out += "NaN"; return; }
  if (std::isinf(v)) { out += (v < 0 ? "-Infinity" : "Infinity"); return; }
  char buf[64];
  auto res = std::to_chars(buf, buf + sizeof(buf), v, std::chars_format::scientific);
  // buf = [-]d[.ddd]e[+-]XX   (shortest round-trip digits)
  const char* p = buf;
  const char* end = res.ptr;
  if (*p == '-') { out += '-'; ++p; }
  char digits[32];
  int nd = 0;
  const char* e = p;
  while (e < end && *e != 'e') {
    if (*e != '.') digits[nd++] = *e;
    ++e;
  }
  int exp10 = 0;
  {
    const char* q = e + 1;
    int sign = 1;
    if (*q == '-') { sign = -1; ++q; } else if (*q == '+') { ++q; }
    while (q < end) exp10 = exp10 * 10 + (*q++ - '0');
    exp10 *= sign;
  }
  while (nd > 1 && digits[nd - 1] == '0') --nd;        // to_chars never pads, but be safe
  const int decpt = exp10 + 1;                         // position of the decimal point
  if (nd == 1 && digits[0] == '0') { out += "0.0"; return; }
  if (decpt <= -4 || decpt > 16) {
    out += digits[0];
    if (nd > 1) { out += '.'; out.append(digits + 1, nd - 1); }
    out += 'e';
    int x = decpt - 1;
    out += (x < 0 ? '-' : '+');
    if (x < 0) x = -x;
    char eb[8];
    int ne = 0;
    do { eb[ne++] = (char)('0' + x % 10); x /= 10; } while (x);
    if (ne < 2) eb[ne++] = '0';
    while (ne) out += eb[--ne];
  } else if (decpt <= 0) {
    out += "0.";
    out.append((size_t)(-decpt), '0');
    out.append(digits, nd);
  } else if (decpt >= nd) {
    out.append(digits, nd);
    out.append((size_t)(decpt - nd), '0');
    out += ".0";
  } else {
    out.append(digits, decpt);
    out += '.';
    out.append(digits + decpt, nd - decpt);
  }
}

inline void put_int(std::string& out, long long v) {
  char buf[24];
  auto res = std::to_chars(buf, buf + sizeof(buf), v);
  out.append(buf, res.ptr - buf);
}

inline void indent(std::string& out, int level) { out.append((size_t)(4 * level), ' '); }

// "key": [ints]  at dict level 2
template <typename T>
void put_int_list(std::string& out, const char* key, const T* v, long long n) {
  indent(out, 2);
  out += '"'; out += key; out += "\": [";
  if (n == 0) { out += ']'; return; }
  for (long long i = 0; i < n; ++i) {
    out += (i ? ",\n" : "\n");
    indent(out, 3);
    put_int(out, (long long)v[i]);
  }
  out += '\n';
  indent(out, 2);
  out += ']';
}

// "key": [[floats] x rows]  at dict level 2
void put_float_matrix(std::string& out, const char* key, const double* v, long long rows, long long cols) {
  indent(out, 2);
  out += '"'; out += key; out += "\": [";
  if (rows == 0) { out += ']'; return; }
  for (long long r = 0; r < rows; ++r) {
    out += (r ? ",\n" : "\n");
    indent(out, 3);
    out += '[';
    if (cols == 0) { out += ']'; continue; }
    for (long long c = 0; c < cols; ++c) {
      out += (c ? ",\n" : "\n");
      indent(out, 4);
      put_float(out, v[r * cols + c]);
    }
    out += '\n';
    indent(out, 3);
    out += ']';
  }
  out += '\n';
  indent(out, 2);
  out += ']';
}

bool flush(FILE* f, std::string& s) {
  const bool ok = fwrite(s.data(), 1, s.size(), f) == s.size();
  s.clear();
  return ok;
}

}  // namespace

// Test hook: float.__repr__ of v into buf (NUL-terminated); returns the length.
extern "C" int mwd_format_float_repr(double v, char* buf, int buf_len) {
  std::string s;
  put_float(s, v);
  if ((int)s.size() + 1 > buf_len) return -1;
  memcpy(buf, s.c_str(), s.size() + 1);
  return (int)s.size();
}

// All arrays are HOST arrays in corpus order.  phone_off / region_off: n_pairs+1 offsets;
// align_probs: per pair (T x n) row-major at ap_off[p];  concept_alignment / concept_probs /
// cluster_probs may be NULL (key omitted).  concept_probs / cluster_probs: (R x n_concepts).
// Key sets written: index, image_concepts, alignment, align_probs, is_phoneme (+ concept_alignment,
// + concept_probs [gaussian :651], + cluster_probs [two-layer]) -- sorted like sort_keys=True.
extern "C" int mwd_write_alignment_files(const char* txt_path, const char* json_path, int64_t n_pairs,
                                         const int64_t* phone_off, const int64_t* region_off,
                                         const int32_t* alignment, const int32_t* image_concepts,
                                         const int32_t* concept_alignment, const double* align_probs,
                                         const int64_t* ap_off, const double* concept_probs,
                                         const double* cluster_probs, int n_concepts, int is_phoneme) {
  FILE* ft = fopen(txt_path, "w");
  if (!ft) { mwd::set_error("cannot open %s", txt_path); return 1; }
  FILE* fj = fopen(json_path, "w");
  if (!fj) { fclose(ft); mwd::set_error("cannot open %s", json_path); return 1; }
  std::string t, j;
  t.reserve(1 << 20);
  j.reserve(1 << 22);
  bool ok = true;
  j += (n_pairs == 0) ? "[]" : "[";
  for (int64_t p = 0; p < n_pairs && ok; ++p) {
    const int64_t T = phone_off[p + 1] - phone_off[p];
    const int64_t n = region_off[p + 1] - region_off[p];
    const int32_t* ali = alignment + phone_off[p];
    for (int64_t i = 0; i < T; ++i) { put_int(t, ali[i]); t += ' '; }
    t += "\n\n";
    j += (p ? ",\n" : "\n");
    indent(j, 1);
    j += "{\n";
    put_float_matrix(j, "align_probs", align_probs + ap_off[p], T, n);
    j += ",\n";
    put_int_list(j, "alignment", ali, T);
    j += ",\n";
    if (cluster_probs) {
      put_float_matrix(j, "cluster_probs", cluster_probs + region_off[p] * n_concepts, n, n_concepts);
      j += ",\n";
    }
    if (concept_alignment) {
      put_int_list(j, "concept_alignment", concept_alignment + phone_off[p], T);
      j += ",\n";
    }
    if (concept_probs) {
      put_float_matrix(j, "concept_probs", concept_probs + region_off[p] * n_concepts, n, n_concepts);
      j += ",\n";
    }
    put_int_list(j, "image_concepts", image_concepts + region_off[p], n);
    j += ",\n";
    indent(j, 2);
    j += "\"index\": ";
    put_int(j, p);
    j += ",\n";
    indent(j, 2);
    j += "\"is_phoneme\": ";
    j += is_phoneme ? "true" : "false";
    j += '\n';
    indent(j, 1);
    j += '}';
    if (j.size() > (1u << 22)) ok = flush(fj, j) && ok;
    if (t.size() > (1u << 20)) ok = flush(ft, t) && ok;
  }
  if (n_pairs) j += "\n]";
  ok = flush(fj, j) && ok;
  ok = flush(ft, t) && ok;
  ok = (fclose(fj) == 0) && ok;
  ok = (fclose(ft) == 0) && ok;
  if (!ok) { mwd::set_error("short write to %s / %s", txt_path, json_path); return 1; }
  return 0;
}
