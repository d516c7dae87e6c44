// On-disk formats of the reference, parsed on the GPU:
//   <id>.csv       the Android log (vmvo/datasets/bdd/bdd_raw.py:53-55: pd.read_csv, column 0 =
//                  Timestamp [ms], then Latitude, Longitude, heading, speed, ...;
//                  vmvo/utils/trajectory.py:191-228 names the columns the path reads)
//   <id>_traj.csv  the cached VO trajectory (bdd_raw.py:150-168, 331-332: columns x, y, z, rot,
//                  rot = str(3x3 ndarray), a quoted field that spans three lines)
// The bytes of any number of files sit concatenated in HBM (every file starts on a 16-byte
// boundary).  Three passes, all HBM-streaming byte work:
//   1. csv_scan_kernel    16 bytes per thread (one coalesced uint4): quote and row-start bit masks,
//                         quote parity by prefix-xor inside the thread, ballot across the warp,
//                         shared memory across the block; per 4 KiB block: parity and the row
//                         counts under both possible entry parities.
//      csv_chain_kernel   one warp per file chains its blocks (ballot / shuffle scan).
//   2. csv_emit_kernel    same masks, now with the true entry parity: row-start byte offsets.
//   3. csv_parse_kernel   one thread per row: split fields (quoted fields, "" escapes), convert the
//                         wanted columns.
// Number conversion is restated from the two converters the reference goes through:
//   * pandas' default (float_precision=None -> precise_xstrtod, pandas/_libs/src/parser/tokenizer.c;
//     third-party, `pandas` unpinned in requirements.txt:9, 3.0.2 in this image): at most 17 digits
//     accumulated in a double, one multiplication or division by a tabulated power of ten.  NOT
//     correctly rounded for 16-17 digit inputs -- reproduced operation by operation, so the result
//     is bit-identical to what the reference reads (tests compare against pandas itself).
//   * the rot entries: np.array(tokens).astype(np.float32) (bdd_raw.py:163-164) = correctly rounded
//     double (Python float()), then rounded to float32.  Here: exact for <= 15 significant digits
//     and |exponent| <= 22 (one IEEE operation on exact operands), double-double product with a
//     106-bit power of ten otherwise.
#include "vmvo_device.cuh"
#include "vmvo_internal.h"

#include <math_constants.h>

namespace vmvo {

#include "vmvo_pow10.inc"

constexpr int kCsvThreads = 256;
constexpr int kCsvChunk = kCsvThreads * 16;   // bytes per block

struct CsvBlk { int parity, cnt0, cnt1, base; };
// after csv_chain_kernel: parity = quote parity on entry to the block, base = rows of the file
// that start before the block

struct CsvMasks { unsigned quote, start; };

// 4-bit mask of the bytes of `w` equal to `c`: exact zero-byte test on w ^ cccc, then the four
// byte-high bits gathered by one multiply (bit 7 of byte i -> bit i)
__device__ __forceinline__ unsigned eq4(unsigned w, unsigned c) {
  const unsigned x = w ^ (c * 0x01010101u);
  const unsigned t = ~((((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) | 0x7f7f7f7fu);   // 0x80 where x byte == 0
  return (((t >> 7) * 0x00204081u) >> 21) & 0xfu;
}

// quote bits and row-start bits of the 16 bytes at `pos` (16-byte aligned) of a file [f0, f1).
// A row starts at byte i when the byte before it is a newline (or i is the first byte of the
// file) and the line is not blank (pandas skip_blank_lines: "\n" or "\r\n" alone); whether that
// newline was inside a quoted field is decided by the caller from the quote parity.
__device__ __forceinline__ CsvMasks csv_masks(const unsigned char* bytes, long long pos, long long f0,
                                              long long f1) {
  CsvMasks m{0u, 0u};
  if (pos >= f1) return m;
  const uint4 q = *reinterpret_cast<const uint4*>(bytes + pos);
  const int n = (f1 - pos) < 16 ? (int)(f1 - pos) : 16;
  const unsigned in = n == 16 ? 0xffffu : (1u << n) - 1u;
  const unsigned quote = eq4(q.x, '"') | eq4(q.y, '"') << 4 | eq4(q.z, '"') << 8 | eq4(q.w, '"') << 12;
  const unsigned nl = (eq4(q.x, '\n') | eq4(q.y, '\n') << 4 | eq4(q.z, '\n') << 8 | eq4(q.w, '\n') << 12) & in;
  const unsigned cr = (eq4(q.x, '\r') | eq4(q.y, '\r') << 4 | eq4(q.z, '\r') << 8 | eq4(q.w, '\r') << 12) & in;
  const unsigned prev_nl = pos > f0 ? bytes[pos - 1] == '\n' : 1u;
  // the byte after the last one of this thread; the end of the file acts as a newline
  const unsigned after_nl = pos + 16 < f1 ? bytes[pos + 16] == '\n' : 1u;
  const unsigned after = nl >> 1 | (n == 16 ? after_nl << 15 : 1u << (n - 1));   // bit i: byte i+1 is a newline
  m.quote = quote & in;
  m.start = (nl << 1 | prev_nl) & in & ~nl & ~(cr & after);
  return m;
}

__device__ __forceinline__ unsigned prefix_xor16(unsigned x) {   // inclusive, 16 bits
  x ^= x << 1; x ^= x << 2; x ^= x << 4; x ^= x << 8;
  return x & 0xffffu;
}

// which file a block belongs to: blk_off[n_files + 1] = prefix of ceil(len / 4096)
__device__ __forceinline__ int file_of_block(const long long* blk_off, int n_files, long long b) {
  int lo = 0, hi = n_files;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (blk_off[mid] <= b) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void csv_block_offsets_kernel(int n_files, const long long* file_off, const long long* file_len,
                                         long long* blk_off) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    long long acc = 0;
    for (int f = 0; f < n_files; ++f) {
      blk_off[f] = acc;
      acc += (file_len[f] + kCsvChunk - 1) / kCsvChunk;
    }
    blk_off[n_files] = acc;
  }
}

// pass 1: the only pass over the bytes.  Per thread: the row starts that count if the block is
// entered outside a quoted field (low half-word) or inside one (high half-word) -- 4 bytes of
// scratch per 16 bytes of file -- and per block the quote parity and both row counts.
__global__ void __launch_bounds__(kCsvThreads)
csv_scan_kernel(const unsigned char* bytes, int n_files, const long long* file_off,
                const long long* file_len, const long long* blk_off, CsvBlk* blk, unsigned* masks) {
  __shared__ int s_par[kCsvThreads / 32];
  __shared__ int s_c0[kCsvThreads / 32], s_c1[kCsvThreads / 32];
  const long long b = blockIdx.x;
  const int f = file_of_block(blk_off, n_files, b);
  const long long f0 = file_off[f], f1 = f0 + file_len[f];
  const long long pos = f0 + (b - blk_off[f]) * kCsvChunk + threadIdx.x * 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const CsvMasks m = csv_masks(bytes, pos, f0, f1);
  const unsigned excl = prefix_xor16(m.quote) ^ m.quote;       // parity before each byte
  const unsigned tp = __popc(m.quote) & 1u;
  const unsigned bal = __ballot_sync(FULL, tp);
  if (lane == 0) s_par[warp] = __popc(bal) & 1;
  __syncthreads();
  int rel = __popc(bal & ((1u << lane) - 1u));
  int blk_par = 0;
#pragma unroll
  for (int q = 0; q < kCsvThreads / 32; ++q) {
    if (q < warp) rel += s_par[q];
    blk_par ^= s_par[q];
  }
  // bit set = inside a quoted field, if the block is entered outside one
  const unsigned inq = excl ^ ((rel & 1) ? 0xffffu : 0u);
  const unsigned v0 = m.start & ~inq, v1 = m.start & inq;
  masks[b * kCsvThreads + threadIdx.x] = v0 | v1 << 16;
  const int c0 = __reduce_add_sync(FULL, __popc(v0));
  const int c1 = __reduce_add_sync(FULL, __popc(v1));
  if (lane == 0) { s_c0[warp] = c0; s_c1[warp] = c1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t0 = 0, t1 = 0;
#pragma unroll
    for (int q = 0; q < kCsvThreads / 32; ++q) { t0 += s_c0[q]; t1 += s_c1[q]; }
    blk[b] = CsvBlk{blk_par, t0, t1, 0};
  }
}

// pass 2: row-start offsets from the stored masks and the chained entry parities
__global__ void __launch_bounds__(kCsvThreads)
csv_emit_kernel(int n_files, const long long* file_off, const long long* blk_off, const CsvBlk* blk,
                const unsigned* masks, const long long* row_off, long long* row_starts) {
  __shared__ int s_cnt[kCsvThreads / 32];
  const long long b = blockIdx.x;
  const int f = file_of_block(blk_off, n_files, b);
  const long long pos = file_off[f] + (b - blk_off[f]) * kCsvChunk + threadIdx.x * 16;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const CsvBlk me = blk[b];
  const unsigned packed = masks[b * kCsvThreads + threadIdx.x];
  const unsigned valid = me.parity ? packed >> 16 : packed & 0xffffu;
  const int cnt = __popc(valid);
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_cnt[warp] = incl;
  __syncthreads();
  long long idx = row_off[f] + me.base + (incl - cnt);
#pragma unroll
  for (int q = 0; q < kCsvThreads / 32; ++q)
    if (q < warp) idx += s_cnt[q];
  unsigned v = valid;
  while (v) {
    const int i = __ffs(v) - 1;
    v &= v - 1;
    row_starts[idx++] = pos + i;
  }
}

// one warp per file: entry parity and row base of each of its blocks, rows of the file
__global__ void csv_chain_kernel(int n_files, const long long* blk_off, CsvBlk* blk, long long* row_counts) {
  const int lane = threadIdx.x & 31;
  const int f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (f >= n_files) return;
  const long long b0 = blk_off[f], b1 = blk_off[f + 1];
  int par = 0;
  long long rows = 0;
  for (long long base = b0; base < b1; base += 32) {
    const long long b = base + lane;
    CsvBlk me{0, 0, 0, 0};
    if (b < b1) me = blk[b];
    const unsigned bal = __ballot_sync(FULL, me.parity & 1);
    const int entry = (par + __popc(bal & ((1u << lane) - 1u))) & 1;
    const int cnt = entry ? me.cnt1 : me.cnt0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += t;
    }
    if (b < b1) {
      me.parity = entry;
      me.base = (int)(rows + incl - cnt);
      blk[b] = me;
    }
    rows += __shfl_sync(FULL, incl, 31);
    par = (par + __popc(bal)) & 1;
  }
  if (lane == 0) row_counts[f] = rows;
}

// ---- number conversion ---------------------------------------------------------------------------
__device__ __forceinline__ bool is_space(unsigned c) { return c == ' ' || (c >= 9 && c <= 13); }
__device__ __forceinline__ bool is_digit(unsigned c) { return c - '0' < 10u; }
__device__ __forceinline__ unsigned lower(unsigned c) { return (c - 'A' < 26u) ? c + 32 : c; }

__device__ bool word_is(const unsigned char* p, int n, const char* w, bool nocase) {
  int i = 0;
  for (; i < n; ++i) {
    if (!w[i]) return false;
    const unsigned a = nocase ? lower(p[i]) : p[i];
    if (a != (unsigned char)w[i]) return false;
  }
  return w[i] == 0;
}

// pandas' default NA strings (pandas/_libs/parsers.pyx STR_NA_VALUES)
__device__ bool is_na_word(const unsigned char* p, int n) {
  if (n == 0) return true;
  if (n > 9) return false;
  return word_is(p, n, "NaN", false) || word_is(p, n, "nan", false) || word_is(p, n, "NA", false) ||
         word_is(p, n, "N/A", false) || word_is(p, n, "n/a", false) || word_is(p, n, "NULL", false) ||
         word_is(p, n, "null", false) || word_is(p, n, "None", false) || word_is(p, n, "<NA>", false) ||
         word_is(p, n, "#N/A", false) || word_is(p, n, "#NA", false) || word_is(p, n, "-NaN", false) ||
         word_is(p, n, "-nan", false) || word_is(p, n, "1.#IND", false) || word_is(p, n, "1.#QNAN", false) ||
         word_is(p, n, "-1.#IND", false) || word_is(p, n, "-1.#QNAN", false) ||
         word_is(p, n, "#N/A N/A", false);
}

// precise_xstrtod, operation by operation.  Returns false when the field is not a number
// (pandas would then make the column `object`).
__device__ bool pandas_float(const unsigned char* p, const unsigned char* end, double* out) {
  const unsigned char* const begin = p;
  while (p < end && is_space(*p)) ++p;
  bool neg = false;
  if (p < end && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
  double number = 0.0;
  int exponent = 0, nd = 0, ndec = 0;
  const int kMaxDigits = 17;
  while (p < end && is_digit(*p)) {
    if (nd < kMaxDigits) { number = dadd(dmul(number, 10.0), (double)(*p - '0')); ++nd; }
    else ++exponent;
    ++p;
  }
  if (p < end && *p == '.') {
    ++p;
    while (nd < kMaxDigits && p < end && is_digit(*p)) {
      number = dadd(dmul(number, 10.0), (double)(*p - '0'));
      ++p; ++nd; ++ndec;
    }
    if (nd >= kMaxDigits) while (p < end && is_digit(*p)) ++p;
    exponent -= ndec;
  }
  bool ok = nd > 0;
  if (ok) {
    if (neg) number = -number;
    if (p < end && (*p == 'e' || *p == 'E')) {
      const unsigned char* save = p;
      ++p;
      bool eneg = false;
      if (p < end && (*p == '-' || *p == '+')) { eneg = *p == '-'; ++p; }
      int n = 0, ed = 0;
      while (ed < kMaxDigits && p < end && is_digit(*p)) { n = n * 10 + (*p - '0'); ++ed; ++p; }
      exponent += eneg ? -n : n;
      if (ed == 0) p = save;   // no digits after the 'e': un-consume it
    }
    if (exponent > 308) number = neg ? -CUDART_INF : CUDART_INF;
    else if (exponent > 0) number = dmul(number, kPow10[exponent]);
    else if (exponent < -308) {
      if (exponent < -616) number = 0.0;
      else { number = ddiv(number, kPow10[-308 - exponent]); number = ddiv(number, kPow10[308]); }
    } else number = ddiv(number, kPow10[-exponent]);
    while (p < end && is_space(*p)) ++p;
    ok = p == end;                       // an overflow is delivered as a signed infinity
  }
  if (ok) { *out = number; return true; }
  // not a plain number: the infinity spellings pandas accepts (parsers.pyx: _try_double), matched
  // against the whole field
  p = begin;
  const unsigned char* q = end;
  double sign = 1.0;
  if (p < q && (*p == '-' || *p == '+')) { sign = *p == '-' ? -1.0 : 1.0; ++p; }
  if (word_is(p, (int)(q - p), "inf", true) || word_is(p, (int)(q - p), "infinity", true)) {
    *out = sign * CUDART_INF;
    return true;
  }
  return false;
}

// correctly rounded double of a decimal token, then float32 (np.array(tokens).astype(np.float32))
__device__ bool numpy_float32(const unsigned char* p, const unsigned char* end, float* out) {
  bool neg = false;
  if (p < end && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
  const int n = (int)(end - p);
  if (word_is(p, n, "nan", true)) { *out = CUDART_NAN_F; return true; }
  if (word_is(p, n, "inf", true) || word_is(p, n, "infinity", true)) {
    *out = neg ? -CUDART_INF_F : CUDART_INF_F;
    return true;
  }
  // w = the digits up to the last non-zero one (trailing zeros carry no information and are
  // folded into the exponent), at most 19 of them; value = w * 10^e10
  unsigned long long w = 0;
  int nd = 0, e10 = 0, tz = 0;
  bool any = false, seen_dot = false, full = false;
  for (; p < end; ++p) {
    if (is_digit(*p)) {
      any = true;
      const unsigned dgt = *p - '0';
      if (full) { if (!seen_dot) ++e10; continue; }      // digits beyond 19 are dropped (numpy never prints them)
      if (seen_dot) --e10;
      if (dgt == 0) { if (w) ++tz; continue; }
      if (nd + tz + 1 > 19) {                            // no room: this digit and the rest are dropped
        full = true;
        ++e10;                 // before the point: one more power of ten; after it: undo the count above
        continue;
      }
      for (; tz > 0; --tz, ++nd) w *= 10ull;
      w = w * 10ull + dgt;
      ++nd;
    } else if (*p == '.' && !seen_dot) seen_dot = true;
    else break;
  }
  if (!any) return false;
  e10 += tz;
  if (p < end && (*p == 'e' || *p == 'E')) {
    ++p;
    bool eneg = false;
    if (p < end && (*p == '-' || *p == '+')) { eneg = *p == '-'; ++p; }
    int k = 0, ed = 0;
    for (; p < end && is_digit(*p); ++p, ++ed) if (k < 100000) k = k * 10 + (*p - '0');
    if (ed == 0) return false;
    e10 += eneg ? -k : k;
  }
  if (p != end) return false;
  double v;
  if (w == 0) v = 0.0;
  else if (e10 + nd > 45) v = CUDART_INF;          // beyond float32 either way
  else if (e10 + nd < -50) v = 0.0;
  else if (w < (1ull << 53) && e10 >= -22 && e10 <= 22) {
    // both operands exact: one correctly rounded IEEE operation
    v = e10 < 0 ? ddiv((double)w, kPow10[-e10]) : dmul((double)w, kPow10[e10]);
  } else {
    // double-double: (wh + wl) * (hi + lo), ~100 bits, then one rounding to double
    const double wh = (double)(w >> 11 << 11), wl = (double)(w & 2047ull);
    int e = e10;
    double scale = 1.0;
    if (e < kPow10ddMin) { scale = kPow10[kPow10ddMin - e]; e = kPow10ddMin; }   // only below 1e-50: flushed to 0 in float32 anyway
    if (e > kPow10ddMax) e = kPow10ddMax;
    const double hi = kPow10dd[e - kPow10ddMin][0], lo = kPow10dd[e - kPow10ddMin][1];
    const double ph = dmul(wh, hi);
    const double pe = fma(wh, hi, -ph);
    const double rest = dadd(dadd(pe, dmul(wh, lo)), dmul(wl, hi));
    v = dadd(ph, rest) / scale;
  }
  const float r = __double2float_rn(v);
  *out = neg ? -r : r;
  return true;
}

// ---- pass 3: one thread per row ------------------------------------------------------------------
constexpr int kCsvMaxCols = 64;
constexpr int kSlotRot = 1000;
constexpr int kParseThreads = 256;
constexpr int kParseStage = 48 * 1024;     // bytes of shared memory a block stages its rows in

// one data row: bytes [pos, end) relative to `src` (the block's shared-memory copy, or the row's
// own start in the file); 32-bit offsets keep the byte walk cheap
__device__ __forceinline__ void csv_parse_row(const unsigned char* __restrict__ src, int pos, int end,
                                              const int* cm, int n_fields, int n_slots, long long n_data,
                                              long long d, double* out, double* rot, int* status) {
  while (end > pos && (src[end - 1] == '\n' || src[end - 1] == '\r')) --end;
  unsigned seen = 0;
  bool seen_rot = false;
  int st = 0, c = 0;
  while (pos <= end) {
    int a, b;                  // content of the field
    if (pos < end && src[pos] == '"') {
      a = ++pos;
      for (;;) {
        if (pos >= end) { b = end; break; }
        if (src[pos] == '"') {
          if (pos + 1 < end && src[pos + 1] == '"') { pos += 2; continue; }
          b = pos++;
          break;
        }
        ++pos;
      }
      while (pos < end && src[pos] != ',') ++pos;
    } else {
      a = pos;
      while (pos < end && src[pos] != ',') ++pos;
      b = pos;
    }
    const int slot = c < kCsvMaxCols ? cm[c] : -1;
    if (slot >= 0 && slot < n_slots) {
      double v = CUDART_NAN;
      if (!is_na_word(src + a, b - a) && !pandas_float(src + a, src + b, &v)) {
        v = CUDART_NAN;
        st |= VMVO_CSV_BAD_NUMBER;
      }
      out[(long long)slot * n_data + d] = v;
      seen |= 1u << slot;
    } else if (slot == kSlotRot && rot) {
      // "[[ a b c]\n [ d e f]\n [ g h i]]": brackets and newlines dropped, split on white space
      int k = 0;
      int t = a;
      bool bad = false;
      while (t < b) {
        while (t < b && (is_space(src[t]) || src[t] == '[' || src[t] == ']')) ++t;
        if (t >= b) break;
        int u = t;
        while (u < b && !is_space(src[u]) && src[u] != '[' && src[u] != ']') ++u;
        float v;
        if (k < 9 && numpy_float32(src + t, src + u, &v)) rot[d * 9 + k] = (double)v;
        else bad = true;
        ++k;
        t = u;
      }
      if (bad || k != 9) {
        st |= VMVO_CSV_BAD_ROT;
        for (int q = 0; q < 9; ++q) rot[d * 9 + q] = CUDART_NAN;
      }
      seen_rot = true;
    }
    ++c;
    if (pos >= end) break;
    ++pos;                      // the comma
    if (pos == end) {           // a trailing comma: one more, empty, field
      const int s2 = c < kCsvMaxCols ? cm[c] : -1;
      if (s2 >= 0 && s2 < n_slots) { out[(long long)s2 * n_data + d] = CUDART_NAN; seen |= 1u << s2; }
      ++c;
      break;
    }
  }
  if (c > n_fields) st |= VMVO_CSV_TOO_MANY_FIELDS;   // pandas: ParserError
  // short rows: pandas pads with NaN
  for (int s = 0; s < n_slots; ++s)
    if (!((seen >> s) & 1u)) out[(long long)s * n_data + d] = CUDART_NAN;
  if (rot && !seen_rot) {
    bool wants = false;
    for (int q = 0; q < kCsvMaxCols; ++q) wants |= cm[q] == kSlotRot;
    if (wants) {
      for (int q = 0; q < 9; ++q) rot[d * 9 + q] = CUDART_NAN;
      st |= VMVO_CSV_BAD_ROT;
    }
  }
  if (st) atomicOr(status, st);
}

// A block takes kParseThreads consecutive rows.  Their bytes are one contiguous span of the buffer:
// it is copied to shared memory with coalesced 16-byte loads, and the threads then walk their rows
// byte by byte out of shared memory (rows start at arbitrary offsets, so walking them in global
// memory costs one L1 tag lookup per thread per byte).  Spans over 48 KiB are read in place.
__global__ void __launch_bounds__(kParseThreads)
csv_parse_kernel(const unsigned char* bytes, int n_files, const long long* file_off,
                 const long long* file_len, const long long* row_off, const long long* row_starts,
                 long long total_rows, const int* colmap, const int* n_fields, int n_slots,
                 long long n_data, double* out, double* rot, int* status) {
  __shared__ uint4 stage[kParseStage / 16];
  const long long buf_end = file_off[n_files - 1] + file_len[n_files - 1];
  for (long long r0 = blockIdx.x * (long long)kParseThreads; r0 < total_rows;
       r0 += (long long)gridDim.x * kParseThreads) {
    const long long r1 = r0 + kParseThreads < total_rows ? r0 + kParseThreads : total_rows;
    const long long lo = row_starts[r0] & ~15LL;
    const long long hi = r1 < total_rows ? row_starts[r1] : buf_end;
    const bool staged = hi - lo <= kParseStage;
    if (staged) {
      const uint4* g = reinterpret_cast<const uint4*>(bytes + lo);
      const int n16 = (int)((hi - lo + 15) >> 4);
      for (int i = threadIdx.x; i < n16; i += kParseThreads) stage[i] = g[i];
    }
    __syncthreads();
    const long long r = r0 + threadIdx.x;
    if (r < r1) {
      int flo = 0, fhi = n_files;     // file of the row
      while (fhi - flo > 1) {
        const int mid = (flo + fhi) >> 1;
        if (row_off[mid] <= r) flo = mid; else fhi = mid;
      }
      const int f = flo;
      if (r != row_off[f]) {          // not the header line
        const long long pos = row_starts[r];
        const long long end = r + 1 < row_off[f + 1] ? row_starts[r + 1] : file_off[f] + file_len[f];
        // data-row index r - (f + 1): one header per earlier file
        if (staged)
          csv_parse_row(reinterpret_cast<const unsigned char*>(stage), (int)(pos - lo), (int)(end - lo),
                        colmap + f * kCsvMaxCols, n_fields[f], n_slots, n_data, r - (f + 1), out, rot,
                        status + f);
        else
          csv_parse_row(bytes + pos, 0, (int)(end - pos), colmap + f * kCsvMaxCols, n_fields[f], n_slots,
                        n_data, r - (f + 1), out, rot, status + f);
      }
    }
    __syncthreads();
  }
}

// the reference sorts the log by Timestamp (bdd_raw.py:55) and then indexes it by LABEL
// (trajectory.py:191-207), which only means something for a log that is already in time order
__global__ void csv_sorted_kernel(int n_files, const long long* row_off, long long n_data,
                                  const double* key, int* status) {
  for (long long d = blockIdx.x * (long long)blockDim.x + threadIdx.x; d < n_data;
       d += (long long)gridDim.x * blockDim.x) {
    // data rows of file f: [row_off[f] - f, row_off[f + 1] - (f + 1))
    int lo = 0, hi = n_files;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (row_off[mid] - mid <= d) lo = mid; else hi = mid;
    }
    if (d > row_off[lo] - lo && key[d] < key[d - 1]) atomicOr(&status[lo], VMVO_CSV_UNSORTED);
  }
}

}  // namespace vmvo

using namespace vmvo;

static long long csv_blocks(int32_t n_files, const int64_t* h_file_len) {
  long long b = 0;
  for (int f = 0; f < n_files; ++f) b += (h_file_len[f] + kCsvChunk - 1) / kCsvChunk;
  return b;
}

extern "C" int64_t vmvo_csv_scratch_bytes(int32_t n_files, const int64_t* h_file_len) {
  if (n_files < 1 || !h_file_len) return 0;
  const long long nb = csv_blocks(n_files, h_file_len);
  return (int64_t)sizeof(long long) * (n_files + 1) + (int64_t)sizeof(CsvBlk) * (nb + 1) +
         (int64_t)sizeof(unsigned) * kCsvThreads * nb;
}

static int csv_check(vmvo_ctx* ctx, const void* d_bytes, int32_t n_files, const int64_t* h_file_off,
                     const int64_t* h_file_len, const void* d_file_off, const void* d_file_len,
                     const void* d_scratch) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (n_files < 1) return fail(ctx, VMVO_ERR_BAD_ARG, "n_files < 1");
  if (!d_bytes || !h_file_off || !h_file_len || !d_file_off || !d_file_len || !d_scratch)
    return fail(ctx, VMVO_ERR_BAD_ARG, "NULL pointer");
  if ((uintptr_t)d_bytes & 15) return fail(ctx, VMVO_ERR_BAD_ARG, "d_bytes must be 16-byte aligned");
  for (int f = 0; f < n_files; ++f) {
    if (h_file_off[f] & 15) return fail(ctx, VMVO_ERR_BAD_ARG, "file %d does not start on a 16-byte boundary", f);
    if (h_file_len[f] < 0) return fail(ctx, VMVO_ERR_BAD_ARG, "file %d has a negative length", f);
  }
  return VMVO_OK;
}

extern "C" int vmvo_csv_count_rows(vmvo_ctx* ctx, const uint8_t* d_bytes, int32_t n_files,
                                   const int64_t* h_file_off, const int64_t* h_file_len,
                                   const int64_t* d_file_off, const int64_t* d_file_len,
                                   void* d_scratch, int64_t* d_row_counts, void* stream) {
  int rc = csv_check(ctx, d_bytes, n_files, h_file_off, h_file_len, d_file_off, d_file_len, d_scratch);
  if (rc) return rc;
  if (!d_row_counts) return fail(ctx, VMVO_ERR_BAD_ARG, "NULL pointer");
  VMVO_ON_DEVICE(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  long long* blk_off = (long long*)d_scratch;
  CsvBlk* blk = (CsvBlk*)(blk_off + n_files + 1);
  const long long nb = csv_blocks(n_files, h_file_len);
  csv_block_offsets_kernel<<<1, 32, 0, st>>>(n_files, (const long long*)d_file_off,
                                             (const long long*)d_file_len, blk_off);
  rc = check_launch(ctx, "csv_block_offsets_kernel");
  if (rc) return rc;
  if (nb > 0) {
    csv_scan_kernel<<<(unsigned)nb, kCsvThreads, 0, st>>>(
        d_bytes, n_files, (const long long*)d_file_off, (const long long*)d_file_len, blk_off, blk,
        (unsigned*)(blk + nb + 1));
    rc = check_launch(ctx, "csv_scan_kernel");
    if (rc) return rc;
  }
  const int warps_per_block = 4;
  csv_chain_kernel<<<(n_files + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(
      n_files, blk_off, blk, (long long*)d_row_counts);
  return check_launch(ctx, "csv_chain_kernel");
}

extern "C" int vmvo_csv_index_rows(vmvo_ctx* ctx, const uint8_t* d_bytes, int32_t n_files,
                                   const int64_t* h_file_off, const int64_t* h_file_len,
                                   const int64_t* d_file_off, const int64_t* d_file_len,
                                   void* d_scratch, const int64_t* d_row_off, int64_t* d_row_starts,
                                   void* stream) {
  int rc = csv_check(ctx, d_bytes, n_files, h_file_off, h_file_len, d_file_off, d_file_len, d_scratch);
  if (rc) return rc;
  if (!d_row_off || !d_row_starts) return fail(ctx, VMVO_ERR_BAD_ARG, "NULL pointer");
  VMVO_ON_DEVICE(ctx);
  long long* blk_off = (long long*)d_scratch;
  CsvBlk* blk = (CsvBlk*)(blk_off + n_files + 1);
  const long long nb = csv_blocks(n_files, h_file_len);
  if (nb == 0) return VMVO_OK;
  csv_emit_kernel<<<(unsigned)nb, kCsvThreads, 0, (cudaStream_t)stream>>>(
      n_files, (const long long*)d_file_off, blk_off, blk, (const unsigned*)(blk + nb + 1),
      (const long long*)d_row_off, (long long*)d_row_starts);
  return check_launch(ctx, "csv_emit_kernel");
}

extern "C" int vmvo_csv_parse_f64(vmvo_ctx* ctx, const uint8_t* d_bytes, int32_t n_files,
                                  const int64_t* d_file_off, const int64_t* d_file_len,
                                  const int64_t* d_row_off, const int64_t* d_row_starts,
                                  int64_t total_rows, const int32_t* d_colmap, const int32_t* d_n_fields,
                                  int32_t n_slots, int32_t sorted_slot, double* d_out, double* d_rot,
                                  int32_t* d_status, void* stream) {
  if (!ctx) return VMVO_ERR_BAD_ARG;
  if (n_files < 1 || total_rows < n_files) return fail(ctx, VMVO_ERR_BAD_ARG, "every file needs a header row");
  if (n_slots < 0 || n_slots > 32) return fail(ctx, VMVO_ERR_BAD_ARG, "n_slots must be in [0, 32]");
  const long long n_data = total_rows - n_files;
  if (!d_bytes || !d_file_off || !d_file_len || !d_row_off || !d_row_starts || !d_colmap || !d_n_fields ||
      !d_status || (n_slots > 0 && n_data > 0 && !d_out))
    return fail(ctx, VMVO_ERR_BAD_ARG, "NULL pointer");
  if (sorted_slot >= n_slots) return fail(ctx, VMVO_ERR_BAD_ARG, "sorted_slot out of range");
  VMVO_ON_DEVICE(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  VMVO_CUDA(ctx, cudaMemsetAsync(d_status, 0, sizeof(int32_t) * n_files, st));
  if (n_data == 0) return VMVO_OK;
  const int threads = kParseThreads;
  long long blocks = (total_rows + threads - 1) / threads;
  if (blocks > ctx->sm_count * 64LL) blocks = ctx->sm_count * 64LL;
  csv_parse_kernel<<<(unsigned)blocks, threads, 0, st>>>(
      d_bytes, n_files, (const long long*)d_file_off, (const long long*)d_file_len,
      (const long long*)d_row_off, (const long long*)d_row_starts, total_rows, d_colmap, d_n_fields,
      n_slots, n_data, d_out, d_rot, d_status);
  int rc = check_launch(ctx, "csv_parse_kernel");
  if (rc) return rc;
  if (sorted_slot >= 0) {
    csv_sorted_kernel<<<(unsigned)blocks, threads, 0, st>>>(n_files, (const long long*)d_row_off, n_data,
                                                           d_out + (long long)sorted_slot * n_data, d_status);
    rc = check_launch(ctx, "csv_sorted_kernel");
  }
  return rc;
}
