// K1 — batched sweep deconvolution (ImpulseResponseEstimator.estimate, reference
// core/impulse_response_estimator.py:149-151).
//
//   out[r] = convolve(rec[r], inverse_filter, mode='same')          for a whole batch of recordings
//
// Two recordings (typically the two ears of one speaker) are packed as re + i*im of ONE complex
// transform: the inverse filter is real, so (a + i b) * h = (a*h) + i (b*h) and no real-FFT
// untangling pass exists.  Only the len(rec) centred samples of the full linear convolution are
// wanted, so the circular length only has to satisfy N >= L + ceil((M-1)/2) (wrap-around lands in
// samples that are discarded); for the 6.15 s / 48 kHz sweep that is 538 905 -> N = 768 x 768 instead
// of SciPy's 691 200-point real transforms.
//
// Three launches per chunk of pairs (chunks sized so the complex workspace stays inside the 126 MB L2):
//   A  column pass: load 2 real rows (zero-padded, any sample format) -> DIF over n1 -> x W_N^(n2 k1) -> ws
//   BC row pass   : ws row -> DIF over n2 -> x H (inverse-sweep spectrum, pre-permuted, pre-scaled)
//                   -> DIT over k2 -> x conj(W_N^(n2 k1)) -> ws           (in place)
//   D  column pass: ws -> DIT over k1 -> write the wanted slice of re / im to the two output rows
// Trips of a radix stage unrolled in pairs in the K1 kernels (fft_fixed.cuh: kTripUnroll): the loads of the second trip are
// issued under the butterflies of the first.  FP64 A / BC / D 3.51 / 4.29 / 2.65 -> 3.36 / 4.24 / 2.63 ms per 1792 IRs
// (BC spills 40 bytes and still gains), FP32 6.76 -> 6.45 ms; unrolling all three trips costs 40 % (registers).
// The frequency-response transforms (fr.cu) keep the rolled loops.
#ifndef IMPB_TRIP_UNROLL
#define IMPB_TRIP_UNROLL 2
#endif
// A/B knob: streaming (evict-first, ld/st.global.cs) accesses for the workspace and output traffic, which is touched
// once per pass, so that it does not displace the 19 MB of spectrum / twiddle tables every ear pair re-reads from L2.
// Measured: pass BC 4.23 -> 4.40 ms per 1792 IRs, A and D unchanged -- off.
#ifndef IMPB_STREAM_HINTS
#define IMPB_STREAM_HINTS 0
#endif
#if IMPB_STREAM_HINTS
#define IMPB_WS_LD(p) __ldcs(p)
#define IMPB_WS_ST(p, v) __stcs((p), (v))
#else
#define IMPB_WS_LD(p) (*(p))
#define IMPB_WS_ST(p, v) (*(p) = (v))
#endif
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <set>
#include <utility>
#include <vector>

#include "fft2d.cuh"
#include "fft_fixed.cuh"

using namespace impb;

namespace {

template <typename T> struct DeconvArgs {
  cx_t<T>* ws;
  const cx_t<T>* tw;
  const cx_t<T>* hs;
  const cx_t<T>* dcol;  // progression steps of the four-step twiddle (Fft2D::dcol / drow)
  const cx_t<T>* drow;
  int64_t N;
  int N1, N2;
  // input
  const void* rec;
  const int64_t* in_offsets;  // optional: start of row r inside rec (else r * ld_in)
  int64_t ld_in, len, nrows, row0;
  int in_fmt;
  // output
  void* out;
  int64_t ld_out, out_start, out_len;
  int out_f64;
  unsigned long long* absmax;  // optional [nrows]: bit pattern of max |out[r, :]| (atomicMax; zeroed by the caller)
  // optional per-(row, column tile) candidates for argmax|out[r, :]|: value bits and output index
  unsigned long long* cand_val;
  int64_t* cand_idx;
  int cand_tiles;
  unsigned long long* blockmax;  // optional [nrows][nblk]: bits of max |out| per time block (specialised path)
  int nblk;
  bool ws_rows_ok(int pairs) const { return (int64_t)pairs * N1 < ((int64_t)1 << 31); }  // TMA coordinates are int32
};

__device__ __forceinline__ double load_sample(const void* p, int fmt, int64_t i) {
  switch (fmt) {
    case IMPB_FMT_F64: return __ldg(reinterpret_cast<const double*>(p) + i);
    case IMPB_FMT_F32: return (double)__ldg(reinterpret_cast<const float*>(p) + i);
    default: return (double)__ldg(reinterpret_cast<const int32_t*>(p) + i) * (1.0 / 2147483648.0);
  }
}

// A: real pair -> column DIF   (PRE: the kernel applies post() itself, store() must not)
template <typename T, int TILE, bool PRE = false> struct DeconvColIn {
  DeconvArgs<T> a;
  __device__ __forceinline__ cx_t<T> load(int b, int t, int e, int c) const {
    const int n2 = t * TILE + c;
    const int64_t n = (int64_t)e * a.N2 + n2;
    if (n2 >= a.N2 || n >= a.len) return mkc<T>(0, 0);
    const int64_t ra = a.row0 + 2 * (int64_t)b, rb = ra + 1;
    const int64_t oa = a.in_offsets ? __ldg(a.in_offsets + ra) : ra * a.ld_in;
    const T re = (T)load_sample(a.rec, a.in_fmt, oa + n);
    T im = (T)0;
    if (rb < a.nrows) {
      const int64_t ob = a.in_offsets ? __ldg(a.in_offsets + rb) : rb * a.ld_in;
      im = (T)load_sample(a.rec, a.in_fmt, ob + n);
    }
    return mkc<T>(re, im);
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, cx_t<T> v) const {
    const int n2 = t * TILE + c;
    if (n2 >= a.N2) return;
    const size_t idx = (size_t)e * a.N2 + n2;
    IMPB_WS_ST(a.ws + (size_t)b * a.N + idx, PRE ? v : cmul(v, __ldg(a.tw + idx)));
  }
#ifndef IMPB_NO_TW_PROGRESSION
  // the specialised kernels fetch post() for output 0 of a butterfly only and multiply up by post_step()^p
  static constexpr bool post_progression = PRE;
#else
  static constexpr bool post_progression = false;
#endif
  __device__ __forceinline__ cx_t<T> post_step(int t, int c) const {
    const int n2 = t * TILE + c;
    return n2 < a.N2 ? __ldg(a.dcol + n2) : mkc<T>(0, 0);
  }
  // four-step twiddle of output (e, c), fetched ahead of the butterfly by the specialised kernels
  __device__ __forceinline__ cx_t<T> post(int t, int e, int c) const {
    const int n2 = t * TILE + c;
#ifdef IMPB_ABL_NO_TABLES
    return mkc<T>((T)1, (T)0);
#else
    return n2 < a.N2 ? __ldg(a.tw + (size_t)e * a.N2 + n2) : mkc<T>(0, 0);
#endif
  }
  // column walk of the staged copy: address of the sample feeding re (which=0) / im (which=1) of element (0, c),
  // the number of leading rows that hold samples (0: the whole column is zero padding) and the row pitch
  __device__ __forceinline__ const T* src_column(int b, int t, int c, int which, int* e_lim, int64_t* e_pitch) const {
    const int n2 = t * TILE + c;
    const int64_t r = a.row0 + 2 * (int64_t)b + which;
    *e_pitch = a.N2;
    if (n2 >= a.N2 || r >= a.nrows || n2 >= a.len) { *e_lim = 0; return reinterpret_cast<const T*>(a.rec); }
    const int64_t o = a.in_offsets ? __ldg(a.in_offsets + r) : r * a.ld_in;
    *e_lim = (int)((a.len - n2 + a.N2 - 1) / a.N2);  // e * N2 + n2 < len
    return reinterpret_cast<const T*>(a.rec) + o + n2;
  }
  // address of the f64 sample feeding re (which=0) / im (which=1) of element (e, c); nullptr = zero padding
  __device__ __forceinline__ const T* src(int b, int t, int e, int c, int which) const {
    const int n2 = t * TILE + c;
    const int64_t n = (int64_t)e * a.N2 + n2;
    const int64_t r = a.row0 + 2 * (int64_t)b + which;
    if (n2 >= a.N2 || n >= a.len || r >= a.nrows) return nullptr;
    const int64_t o = a.in_offsets ? __ldg(a.in_offsets + r) : r * a.ld_in;
    return reinterpret_cast<const T*>(a.rec) + o + n;
  }
};

// BC: row DIF x H x row DIT, in place
template <typename T, int TILE, bool PRE = false> struct DeconvRow {
  DeconvArgs<T> a;
  __device__ __forceinline__ cx_t<T> load(int b, int t, int e, int c) const {
    const int row = t * TILE + c;
    if (row >= a.N1) return mkc<T>(0, 0);
    return IMPB_WS_LD(a.ws + (size_t)b * a.N + (size_t)row * a.N2 + e);
  }
  __device__ __forceinline__ cx_t<T> pw(int t, int R, int items, int q, int w) const {
#ifdef IMPB_ABL_NO_TABLES  // developer ablation (tools/build_variant.sh): what the table loads cost
    return mkc<T>((T)1, (T)0);
#else
    return __ldg(a.hs + ((size_t)t * R + q) * items + w);
#endif
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, cx_t<T> v) const {
    const int row = t * TILE + c;
    if (row >= a.N1) return;
    const size_t idx = (size_t)row * a.N2 + e;
    IMPB_WS_ST(a.ws + (size_t)b * a.N + idx, PRE ? v : cmul(v, cconj(__ldg(a.tw + idx))));
  }
#ifndef IMPB_NO_TW_PROGRESSION
  static constexpr bool post_progression = PRE;
#else
  static constexpr bool post_progression = false;
#endif
  __device__ __forceinline__ cx_t<T> post_step(int t, int c) const {
    const int row = t * TILE + c;
    return row < a.N1 ? cconj(__ldg(a.drow + row)) : mkc<T>(0, 0);
  }
  __device__ __forceinline__ cx_t<T> post(int t, int e, int c) const {
    const int row = t * TILE + c;
#ifdef IMPB_ABL_NO_TABLES
    return mkc<T>((T)1, (T)0);
#else
    return row < a.N1 ? cconj(__ldg(a.tw + (size_t)row * a.N2 + e)) : mkc<T>(0, 0);
#endif
  }
};

// D: column DIT -> real pair out
template <typename T, int TILE> struct DeconvColOut {
  DeconvArgs<T> a;
  __device__ __forceinline__ cx_t<T> load(int b, int t, int e, int c) const {
    const int n2 = t * TILE + c;
    if (n2 >= a.N2) return mkc<T>(0, 0);
    return IMPB_WS_LD(a.ws + (size_t)b * a.N + (size_t)e * a.N2 + n2);
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, cx_t<T> v) const {
    T ax, ay;
    int j;
    store_abs(b, t, e, c, v, ax, ay, j);
  }
  // stores and reports |value as stored| of both rows and the output index; false if the sample lies outside the
  // requested output window
  __device__ __forceinline__ bool store_abs(int b, int t, int e, int c, cx_t<T> v, T& ax, T& ay, int& jj) const {
    const int n2 = t * TILE + c;
    const int64_t j = (int64_t)e * a.N2 + n2 - a.out_start;
    if (n2 >= a.N2 || j < 0 || j >= a.out_len) return false;
    const int64_t ra = a.row0 + 2 * (int64_t)b, rb = ra + 1;
    if (a.out_f64) {
      double* o = reinterpret_cast<double*>(a.out);
      IMPB_WS_ST(o + ra * a.ld_out + j, (double)v.x);
      if (rb < a.nrows) IMPB_WS_ST(o + rb * a.ld_out + j, (double)v.y);
    } else {
      float* o = reinterpret_cast<float*>(a.out);
      v.x = (T)(float)v.x;
      v.y = (T)(float)v.y;
      IMPB_WS_ST(o + ra * a.ld_out + j, (float)v.x);
      if (rb < a.nrows) IMPB_WS_ST(o + rb * a.ld_out + j, (float)v.y);
    }
    ax = fabs(v.x);
    ay = fabs(v.y);
    jj = (int)j;
    return true;
  }
  __device__ __forceinline__ bool tracks_blocks() const { return a.blockmax != nullptr; }
  __device__ __forceinline__ void publish_block(int b, int p, int which, T m) const {
    const int64_t r = a.row0 + 2 * (int64_t)b + which;
    if (r >= a.nrows || !(m > (T)0)) return;
    unsigned long long* dst = a.blockmax + r * a.nblk + p;
    const unsigned long long bits = (unsigned long long)__double_as_longlong((double)m);
    if (bits > __ldcg(dst)) atomicMax(dst, bits);
  }
  __device__ __forceinline__ bool tracks_max() const { return a.absmax != nullptr; }
  __device__ __forceinline__ bool tracks_arg() const { return a.cand_val != nullptr; }
  __device__ __forceinline__ void publish_arg(int b, int t, T mx_re, int ix_re, T mx_im, int ix_im) const {
    const int64_t ra = a.row0 + 2 * (int64_t)b, rb = ra + 1;
    a.cand_val[ra * a.cand_tiles + t] = (unsigned long long)__double_as_longlong((double)mx_re);
    a.cand_idx[ra * a.cand_tiles + t] = ix_re;
    if (rb < a.nrows) {
      a.cand_val[rb * a.cand_tiles + t] = (unsigned long long)__double_as_longlong((double)mx_im);
      a.cand_idx[rb * a.cand_tiles + t] = ix_im;
    }
  }
  __device__ __forceinline__ void publish_max(int b, T mx_re, T mx_im) const {
    const int64_t ra = a.row0 + 2 * (int64_t)b, rb = ra + 1;
    const unsigned long long br = (unsigned long long)__double_as_longlong((double)mx_re);
    const unsigned long long bi = (unsigned long long)__double_as_longlong((double)mx_im);
    if (br > __ldcg(a.absmax + ra)) atomicMax(a.absmax + ra, br);
    if (rb < a.nrows && bi > __ldcg(a.absmax + rb)) atomicMax(a.absmax + rb, bi);
  }
};

// max |row| of an already written output (generic-engine plans, where the column pass does not track it)
template <typename O>
__global__ void __launch_bounds__(256) k_row_absmax(const O* __restrict__ out, int64_t ld, int64_t n,
                                                    unsigned long long* __restrict__ absmax) {
  const int64_t r = blockIdx.x;
  const O* p = out + r * ld;
  int64_t C = (n + gridDim.y - 1) / gridDim.y;
  const int64_t lo = (int64_t)blockIdx.y * C, hi = min(n, lo + C);
  double m = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) m = fmax(m, fabs((double)p[i]));
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(absmax + r, (unsigned long long)__double_as_longlong(m));
}

// argmax from the per-tile candidates of the specialised final pass: smallest index whose value equals the row max
__global__ void __launch_bounds__(256)
k_argmax_pick(const unsigned long long* __restrict__ absmax, const unsigned long long* __restrict__ cand_val,
              const int64_t* __restrict__ cand_idx, int tiles, int64_t* __restrict__ argmax, int64_t rows) {
  const int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);  // one warp per row
  if (r >= rows) return;
  const unsigned long long m = absmax[r];
  unsigned best = 0xffffffffu;
  for (int t = threadIdx.x & 31; t < tiles; t += 32)
    if (cand_val[r * tiles + t] == m) best = min(best, (unsigned)cand_idx[r * tiles + t]);
  best = __reduce_min_sync(0xffffffffu, best);
  if ((threadIdx.x & 31) == 0) argmax[r] = best == 0xffffffffu ? -1 : (int64_t)best;
}

// generic-engine plans: argmax by one more read of the rows (absmax is final when this runs)
template <typename O>
__global__ void __launch_bounds__(256) k_row_argmax(const O* __restrict__ out, int64_t ld, int64_t n,
                                                    const unsigned long long* __restrict__ absmax,
                                                    unsigned long long* __restrict__ argmax) {
  const int64_t r = blockIdx.x;
  const O* p = out + r * ld;
  const double m = __longlong_as_double((long long)absmax[r]);
  int64_t C = (n + gridDim.y - 1) / gridDim.y;
  const int64_t lo = (int64_t)blockIdx.y * C, hi = min(n, lo + C);
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256)
    if (fabs((double)p[i]) == m) atomicMin(argmax + r, (unsigned long long)i);
}

// spectrum of the inverse filter: scrambled [p1][p2] -> natural order
__global__ void k_spec_to_natural(const double2* __restrict__ ws, double2* __restrict__ nat, const int* __restrict__ perm1,
                                  const int* __restrict__ perm2, int N1, int N2) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)N1 * N2) return;
  const int p1 = (int)(i / N2), p2 = (int)(i % N2);
  nat[(size_t)perm1[p1] + (size_t)N1 * perm2[p2]] = ws[i];
}

// natural order -> the register/lane order of the fused row kernel:  hs[(tile*R + q)*items + w]
template <typename T>
__global__ void k_spec_layout(const double2* __restrict__ nat, cx_t<T>* __restrict__ hs, const int* __restrict__ perm1,
                              const int* __restrict__ perm2, int N1, int N2, int TILE, int R, double scale,
                              int j_fastest) {
  const int items = (N2 / R) * TILE;
  const int64_t total = (int64_t)((N1 + TILE - 1) / TILE) * R * items;
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int w = (int)(i % items);
  const int q = (int)((i / items) % R);
  const int t = (int)(i / ((int64_t)items * R));
  // work-item order of the consuming kernel: specialised row pass = j fastest, generic = c fastest
  const int c = j_fastest ? w / (N2 / R) : w % TILE, j = j_fastest ? w % (N2 / R) : w / TILE;
  const int row = t * TILE + c, p2 = j * R + q;
  double2 v = make_double2(0.0, 0.0);
  if (row < N1) v = nat[(size_t)perm1[row] + (size_t)N1 * perm2[p2]];
  hs[i] = mkc<T>((T)(v.x * scale), (T)(v.y * scale));
}

// ---- TMA descriptor of the K1 workspace: [pairs * N1 rows][2 * N2 doubles], box = 256 rows x 128 bytes, 128-byte swizzle ----
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
  static tmap_encode_fn fn = []() -> tmap_encode_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return (tmap_encode_fn)p;
  }();
  return fn;
}
static bool ws_column_map(CUtensorMap* map, const double2* ws, int64_t rows, int N2) {
  tmap_encode_fn enc = tmap_encoder();
  if (!enc) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)2 * N2, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)2 * N2 * sizeof(double)};
  const cuuint32_t box[2] = {16, 256};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)ws, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- compile-time specialised launches (fft_fixed.cuh); sizes must match plan.cu:fixed_radices ----
template <class K> static int fx_prepare(K kernel, size_t smem, int device) {
  // the opt-in shared-memory size is a per-kernel, per-device attribute: set it once
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  const std::pair<const void*, int> key((const void*)kernel, device);
  std::lock_guard<std::mutex> lk(mu);
  if (done.count(key)) return IMPB_OK;
  IMPB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  done.insert(key);
  return IMPB_OK;
}

// threads per CTA of the half-width instantiations (A/B: 128 = four 49 KB CTAs per SM instead of two 256-thread ones)
#ifndef IMPB_HALF_NT
#define IMPB_HALF_NT 256
#endif
// WIDTHS_: 1 = full-width tiles only, 2 = half-width only (TileOf<T>/2 columns: 88.2 ... 192 kHz plans, whose
// sub-FFTs do not fit shared memory twice per SM at full width), 3 = both; a kernel is only instantiated for the tile
// widths its length is planned with
#define IMPB_FX_CASE(KERNEL, S_, R0_, R1_, R2_, WIDTHS_)                                        \
  case S_: {                                                                                    \
    if constexpr ((TILE == TileOf<T>::value && ((WIDTHS_) & 1)) || (TILE == TileOf<T>::value / 2 && ((WIDTHS_) & 2))) { \
      constexpr int NTX = (TILE == TileOf<T>::value) ? 256 : IMPB_HALF_NT;                        \
      auto k = KERNEL<T, TILE, NTX, Radix3<R0_, R1_, R2_>, F>;                                   \
      const size_t smem = fx_tile_elems<S_, TILE>() * sizeof(cx_t<T>);                           \
      int rc = fx_prepare(k, smem, device);                                                      \
      if (rc) return rc;                                                                         \
      k<<<grid, NTX, smem, st>>>(tw0, tw1, f);                                                   \
      IMPB_CUDA(cudaGetLastError());                                                             \
      return IMPB_OK;                                                                            \
    } else {                                                                                     \
      set_error("specialised kernel for S=%d is built for another tile width", S);              \
      return IMPB_EUNSUPPORTED;                                                                  \
    }                                                                                            \
  }

#define IMPB_FX_DISPATCH(NAME, KERNEL)                                                                        \
  template <typename T, int TILE, class F>                                                                   \
  static int NAME(int S, dim3 grid, cudaStream_t st, int device, const cx_t<T>* tw0, const cx_t<T>* tw1,     \
                  const F& f) {                                                                              \
    switch (S) {                                                                                             \
      IMPB_FX_CASE(KERNEL, 768, 8, 12, 8, 3)                                                                 \
      IMPB_FX_CASE(KERNEL, 1536, 12, 8, 16, 2)                                                               \
      IMPB_FX_CASE(KERNEL, 1600, 16, 10, 10, 2)                                                              \
      default: set_error("no specialised kernel for S=%d", S); return IMPB_EUNSUPPORTED;                    \
    }                                                                                                        \
  }
IMPB_FX_DISPATCH(fx_fwd, k_fx_fwd)
IMPB_FX_DISPATCH(fx_fwd_async, k_fx_fwd_async)
IMPB_FX_DISPATCH(fx_inv, k_fx_inv)
IMPB_FX_DISPATCH(fx_conv, k_fx_conv)

template <typename T> struct DeconvEngine {
  Fft2D<T> fft;
  cx_t<T>* hs = nullptr;
  int device = 0;
  static constexpr int FULL = TileOf<T>::value;  // 8 (f64) / 16 (f32)

  // lengths with an instantiated K1 kernel triple (IMPB_FX_DISPATCH below); other specialised lengths of the menu
  // in plan.cu:fixed_radices only serve the frequency-response transforms
  bool fixed_path() const {
    const Plan2D& p = fft.p;
    auto k1_half = [](int S) { return S == 768 || S == 1536 || S == 1600; };
    const bool full = p.tile == FULL && p.N1 == 768 && p.N2 == 768;          // 44.1 / 48 kHz sweeps
    const bool half = p.tile == FULL / 2 && k1_half(p.N1) && k1_half(p.N2);  // 88.2 ... 192 kHz sweeps
    return fft.col_fixed && fft.row_fixed && (full || half) && getenv("IMPB_NO_FIXED") == nullptr;
  }

  int launch(const DeconvArgs<T>& a, int pairs, cudaStream_t st) const {
    return fft.p.tile == FULL ? launch_t<FULL>(a, pairs, st) : launch_t<FULL / 2>(a, pairs, st);
  }

  template <int TILE> int launch_t(const DeconvArgs<T>& a, int pairs, cudaStream_t st) const {
    int rc;
    static const bool no_fixed = getenv("IMPB_NO_FIXED") != nullptr;
    if (TILE == fft.p.tile && fixed_path() && !no_fixed) {
      const Plan2D& p = fft.p;
      static const bool no_async = getenv("IMPB_NO_ASYNC") != nullptr;
      // the staged copy moves raw samples: the input must already be in the tile's arithmetic type
      if (a.in_fmt == (sizeof(T) == 8 ? IMPB_FMT_F64 : IMPB_FMT_F32) && !no_async) {
        rc = fx_fwd_async<T, TILE>(p.N1, dim3(fft.col_tiles(), pairs), st, device, fft.fx_col[0], fft.fx_col[1],
                                   DeconvColIn<T, TILE, true>{a});
      } else {
        rc = fx_fwd<T, TILE>(p.N1, dim3(fft.col_tiles(), pairs), st, device, fft.fx_col[0], fft.fx_col[1],
                             DeconvColIn<T, TILE, true>{a});
      }
      if (rc) return rc;
      if ((rc = fx_conv<T, TILE>(p.N2, dim3(fft.row_tiles(), pairs), st, device, fft.fx_row[0], fft.fx_row[1],
                                 DeconvRow<T, TILE, true>{a})))
        return rc;
      if constexpr (sizeof(T) == 8 && TILE == 8) {
        // TMA-fed column DIT (768-row tiles): the workspace tile arrives through a tensor map, swizzled in flight
        static const bool no_tma = getenv("IMPB_NO_TMA") != nullptr;
        CUtensorMap map;
        if (!no_tma && p.N1 == 768 && a.ws_rows_ok(pairs) && ws_column_map(&map, a.ws, (int64_t)pairs * p.N1, p.N2)) {
          auto k = k_fx_inv_tma<TILE, 256, Radix3<8, 12, 8>, DeconvColOut<T, TILE>>;
          const size_t smem = fx_tile_elems<768, TILE>() * sizeof(cx_t<T>) + 1024;  // + room for the 1024-byte alignment
          if ((rc = fx_prepare(k, smem, device))) return rc;
          k<<<dim3(fft.col_tiles(), pairs), 256, smem, st>>>(fft.fx_col[0], fft.fx_col[1], DeconvColOut<T, TILE>{a}, map);
          IMPB_CUDA(cudaGetLastError());
          return IMPB_OK;
        }
      }
      return fx_inv<T, TILE>(p.N1, dim3(fft.col_tiles(), pairs), st, device, fft.fx_col[0], fft.fx_col[1],
                             DeconvColOut<T, TILE>{a});
    }
    if ((rc = fft.template col_fwd<TILE, false>(DeconvColIn<T, TILE>{a}, pairs, st))) return rc;
    if ((rc = fft.template row_conv<TILE>(DeconvRow<T, TILE>{a}, pairs, st))) return rc;
    if ((rc = fft.template col_inv<TILE, false>(DeconvColOut<T, TILE>{a}, pairs, st))) return rc;
    if (a.absmax) {  // the generic column pass does not track maxima: one extra read of the rows just written
      const int64_t r0 = a.row0, nr = std::min<int64_t>(2 * (int64_t)pairs, a.nrows - r0);
      const dim3 grid((unsigned)nr, (unsigned)std::max<int64_t>(1, std::min<int64_t>(a.out_len / 4096, 1024)));
      if (a.out_f64)
        k_row_absmax<double><<<grid, 256, 0, st>>>((const double*)a.out + r0 * a.ld_out, a.ld_out, a.out_len, a.absmax + r0);
      else
        k_row_absmax<float><<<grid, 256, 0, st>>>((const float*)a.out + r0 * a.ld_out, a.ld_out, a.out_len, a.absmax + r0);
      IMPB_CUDA(cudaGetLastError());
    }
    return IMPB_OK;
  }
};

}  // namespace

struct impb_plan {
  impb_ctx* ctx = nullptr;
  int precision = 64;
  int64_t M = 0, L = 0, s = 0;  // s = (M-1)/2: start of the 'same' slice in the full convolution
  DeconvEngine<double> e64;
  DeconvEngine<float> e32;
  int64_t chunk_pairs = 1;
  bool has64 = false, has32 = false;
  // optional intra-call concurrency: chunks are issued round-robin on `nstreams` internal streams (each with its
  // own workspace slice) so that the tail of one launch triple overlaps the head of the next
  int nstreams = 1;
  cudaStream_t streams[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t fork_ev = nullptr;
  cudaEvent_t join_ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

static int build_spectrum(impb_plan* pl, const double* inv_host) {
  // FP64 transform of the zero-padded filter with a double plan of the same (N1, N2)
  const Plan2D& target = pl->precision == 64 ? pl->e64.fft.p : pl->e32.fft.p;
  const int N1 = target.N1, N2 = target.N2;
  const size_t N = (size_t)target.N;
  Fft2D<double> tmp;
  Fft2D<double>* f64 = nullptr;
  if (pl->precision == 64) {
    f64 = &pl->e64.fft;
  } else {
    Plan2D pd = target;
    pd.tile = target.tile / 2;  // the FP32 plan was derived from an FP64 plan of half its tile width
    pd.threads = 512;
    if (!make_subplan(N1, pd.tile, 512, &pd.col) || !make_subplan(N2, pd.tile, 512, &pd.row)) return IMPB_EUNSUPPORTED;
    pd.smem_col = tile_smem_bytes<double>(pd.col, pd.tile);
    pd.smem_row = tile_smem_bytes<double>(pd.row, pd.tile);
    if (std::max(pd.smem_col, pd.smem_row) > pl->ctx->smem_optin) {
      set_error("FP32 plan %dx%d does not fit the FP64 spectrum builder", N1, N2);
      return IMPB_EUNSUPPORTED;
    }
    int rc = tmp.init(pd);
    if (rc) return rc;
    f64 = &tmp;
  }
  double *d_h = nullptr;
  double2 *ws = nullptr, *nat = nullptr;
  IMPB_CUDA(cudaMalloc(&d_h, pl->M * sizeof(double)));
  IMPB_CUDA(cudaMalloc(&ws, N * sizeof(double2)));
  IMPB_CUDA(cudaMalloc(&nat, N * sizeof(double2)));
  IMPB_CUDA(cudaMemcpy(d_h, inv_host, pl->M * sizeof(double), cudaMemcpyHostToDevice));
  DeconvArgs<double> a{};
  a.ws = ws; a.tw = f64->tw; a.hs = nullptr; a.N = (int64_t)N; a.N1 = N1; a.N2 = N2;
  a.rec = d_h; a.ld_in = pl->M; a.len = pl->M; a.nrows = 1; a.row0 = 0; a.in_fmt = IMPB_FMT_F64;
  int rc;
  if (f64->p.tile == 8) {
    rc = f64->col_fwd<8, false>(DeconvColIn<double, 8>{a}, 1, 0);
    if (!rc) rc = f64->row_fwd<8, false>(RowPlain<double, 8>{ws, ws, nullptr, (int64_t)N, N1, N2, 0, 0}, 1, 0);
  } else {
    rc = f64->col_fwd<4, false>(DeconvColIn<double, 4>{a}, 1, 0);
    if (!rc) rc = f64->row_fwd<4, false>(RowPlain<double, 4>{ws, ws, nullptr, (int64_t)N, N1, N2, 0, 0}, 1, 0);
  }
  if (!rc) {
    const int thr = 256;
    k_spec_to_natural<<<(unsigned)((N + thr - 1) / thr), thr>>>(ws, nat, f64->perm_col, f64->perm_row, N1, N2);
    const double scale = 1.0 / (double)N;
    if (pl->precision == 64) {
      const int R = target.row.radix[target.row.nst - 1];
      const int tl = target.tile;
      const size_t total = (size_t)((N1 + tl - 1) / tl) * tl * N2;
      IMPB_CUDA(cudaMalloc(&pl->e64.hs, total * sizeof(double2)));
      k_spec_layout<double><<<(unsigned)((total + thr - 1) / thr), thr>>>(nat, pl->e64.hs, pl->e64.fft.perm_col,
                                                                       pl->e64.fft.perm_row, N1, N2, tl, R, scale,
                                                                       1);
    } else {
      const int R = target.row.radix[target.row.nst - 1];
      const int tl = target.tile;
      const size_t total = (size_t)((N1 + tl - 1) / tl) * tl * N2;
      IMPB_CUDA(cudaMalloc(&pl->e32.hs, total * sizeof(float2)));
      k_spec_layout<float><<<(unsigned)((total + thr - 1) / thr), thr>>>(nat, pl->e32.hs, pl->e32.fft.perm_col,
                                                                      pl->e32.fft.perm_row, N1, N2, tl, R, scale,
                                                                      1);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = cuda_fail(e, "inverse-filter spectrum build");
  }
  cudaFree(d_h); cudaFree(ws); cudaFree(nat);
  if (pl->precision != 64) tmp.release();
  return rc;
}

extern "C" {

int32_t impb_deconv_plan_create(impb_ctx* ctx, const double* inv_filter_host, int64_t M, int64_t L,
                                int32_t precision, impb_plan** out) {
  if (!ctx || !inv_filter_host || !out || M < 1 || L < 1 || (precision != 64 && precision != 32)) {
    set_error("impb_deconv_plan_create: bad argument");
    return IMPB_EINVAL;
  }
  DeviceGuard g(ctx->device);
  impb_plan* pl = new impb_plan();
  pl->ctx = ctx;
  pl->e64.device = pl->e32.device = ctx->device;
  pl->precision = precision;
  pl->M = M;
  pl->L = L;
  pl->s = (M - 1) / 2;
  // circular length: wanted full-convolution indices [s, s+L) must not be hit by wrap-around
  const int64_t n_min = std::max<int64_t>(std::max(L + (M - 1) - pl->s, pl->s + L), 16);
  Plan2D p;
  // both precisions are planned against the FP64 shared-memory limit (the spectrum is built in FP64)
  if (!plan_2d(ctx, n_min, false, 8, &p, /*allow_half_tile=*/true)) {
    set_error("no 2-pass FFT plan for n >= %lld", (long long)n_min);
    delete pl;
    return IMPB_EUNSUPPORTED;
  }
  int rc;
  if (precision == 64) {
    rc = pl->e64.fft.init(p);
    pl->has64 = true;
  } else {
    Plan2D pf = p;
    // FP32 tiles: 16 columns (128-byte rows, 98 KB, 2 CTAs per SM) by default; IMPB_F32_TILE8 plans 8 columns
    // (64-byte rows, 49 KB: registers then allow 3 CTAs per SM) for A/B measurements
    static const bool f32_narrow = getenv("IMPB_F32_TILE8") != nullptr;
    pf.tile = f32_narrow ? p.tile : 2 * p.tile;
    pf.threads = p.threads;
    if (!make_subplan(p.N1, pf.tile, pf.threads, &pf.col) || !make_subplan(p.N2, pf.tile, pf.threads, &pf.row)) {
      delete pl;
      return IMPB_EUNSUPPORTED;
    }
    pf.smem_col = tile_smem_bytes<float>(pf.col, pf.tile);
    pf.smem_row = tile_smem_bytes<float>(pf.row, pf.tile);
    rc = pl->e32.fft.init(pf);
    pl->has32 = true;
  }
  if (!rc) rc = build_spectrum(pl, inv_filter_host);
  if (rc) {
    impb_deconv_plan_destroy(pl);
    return rc;
  }
  const size_t cx = precision == 64 ? sizeof(double2) : sizeof(float2);
  // pairs per launch triple: measured on B200, fewer/larger launches beat L2-resident chunks at the
  // current kernel speed (tools/quick_bench.py, IMPB_CHUNK_PAIRS sweep: 6 pairs 3.53 ms, 37 pairs 3.00 ms, 113 pairs
  // 2.80 ms, 224 pairs 2.75 ms per 448 IRs); cap the workspace at 4 GiB of the 180 GB
  int64_t chunk = (int64_t)((size_t)4 << 30) / (int64_t)(p.N * cx);
  if (const char* e = getenv("IMPB_CHUNK_PAIRS")) chunk = atoll(e);
  pl->chunk_pairs = std::max<int64_t>(1, std::min<int64_t>(chunk, 32768));
  if (const char* e = getenv("IMPB_STREAMS")) pl->nstreams = std::max(1, std::min(4, atoi(e)));
  if (pl->nstreams > 1) {
    cudaEventCreateWithFlags(&pl->fork_ev, cudaEventDisableTiming);
    for (int i = 0; i < pl->nstreams; ++i) {
      cudaStreamCreateWithFlags(&pl->streams[i], cudaStreamNonBlocking);
      cudaEventCreateWithFlags(&pl->join_ev[i], cudaEventDisableTiming);
    }
  }
  *out = pl;
  return IMPB_OK;
}

void impb_deconv_plan_destroy(impb_plan* pl) {
  if (!pl) return;
  DeviceGuard g(pl->ctx->device);
  for (int i = 0; i < 4; ++i) {
    if (pl->streams[i]) cudaStreamDestroy(pl->streams[i]);
    if (pl->join_ev[i]) cudaEventDestroy(pl->join_ev[i]);
  }
  if (pl->fork_ev) cudaEventDestroy(pl->fork_ev);
  if (pl->has64) { pl->e64.fft.release(); cudaFree(pl->e64.hs); }
  if (pl->has32) { pl->e32.fft.release(); cudaFree(pl->e32.hs); }
  delete pl;
}

int32_t impb_deconv_plan_info(const impb_plan* pl, int64_t* nfft, int64_t* n1, int64_t* n2) {
  if (!pl) return IMPB_EINVAL;
  const Plan2D& p = pl->precision == 64 ? pl->e64.fft.p : pl->e32.fft.p;
  if (nfft) *nfft = p.N;
  if (n1) *n1 = p.N1;
  if (n2) *n2 = p.N2;
  return IMPB_OK;
}

int64_t impb_deconv_workspace_bytes(const impb_plan* pl, int64_t batch) {
  if (!pl || batch < 1) return 0;
  const Plan2D& p = pl->precision == 64 ? pl->e64.fft.p : pl->e32.fft.p;
  const int64_t pairs = std::min<int64_t>((batch + 1) / 2, pl->chunk_pairs);
  return pl->nstreams * pairs * p.N * (int64_t)(pl->precision == 64 ? sizeof(double2) : sizeof(float2));
}

int32_t impb_deconv_launches_per_execute(const impb_plan* pl, int64_t batch) {
  if (!pl || batch < 1) return 0;
  const int64_t pairs = (batch + 1) / 2;
  return (int32_t)(3 * ((pairs + pl->chunk_pairs - 1) / pl->chunk_pairs));
}

int32_t impb_deconv_blockmax_info(const impb_plan* pl, int64_t out_offset, int64_t* nblk, int64_t* blk_len,
                                  int64_t* blk_first) {
  if (!pl || !nblk || !blk_len || !blk_first) return IMPB_EINVAL;
  const bool fixed = pl->precision == 64 ? pl->e64.fixed_path() : pl->e32.fixed_path();
  const Plan2D& p = pl->precision == 64 ? pl->e64.fft.p : pl->e32.fft.p;
  const int r0 = p.col.radix[0];  // outputs of the last DIT stage of the column pass = time blocks of the row
  *nblk = fixed ? r0 : 0;
  *blk_len = fixed ? (int64_t)(p.N1 / r0) * p.N2 : 0;
  *blk_first = -(pl->s + out_offset);
  return IMPB_OK;
}

int32_t impb_deconv_execute(impb_plan* pl, const void* rec, const int64_t* in_offsets, int64_t ld_in, int32_t in_fmt,
                            int64_t len, void* out, int64_t ld_out, int32_t out_f64, int64_t out_offset, int64_t out_len,
                            double* absmax_out, int64_t* argmax_out, double* blockmax_out, int64_t batch, void* workspace,
                            void* stream) {
  if (!pl || !rec || !out || !workspace || batch < 0 || len < 0 || len > pl->L || (!in_offsets && ld_in < len) ||
      out_offset < 0 || (argmax_out && !absmax_out) ||
      out_len < 0 || out_offset + out_len > len || ld_out < out_len || in_fmt < 0 || in_fmt > 2) {
    set_error("impb_deconv_execute: bad argument (len=%lld L=%lld out_offset=%lld out_len=%lld)", (long long)len,
              pl ? (long long)pl->L : 0LL, (long long)out_offset, (long long)out_len);
    return IMPB_EINVAL;
  }
  if (batch == 0 || out_len == 0) return IMPB_OK;
  DeviceGuard g(pl->ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (absmax_out) IMPB_CUDA(cudaMemsetAsync(absmax_out, 0, sizeof(double) * batch, st));
  // argmax: per-(row, tile) candidates from the specialised final pass, or one extra read for generic plans
  const bool fixed = pl->precision == 64 ? pl->e64.fixed_path() : pl->e32.fixed_path();
  // one argmax candidate per (row, column tile, warp of the 256-thread CTA)
  const int cand_tiles = 8 * (pl->precision == 64 ? pl->e64.fft.col_tiles() : pl->e32.fft.col_tiles());
  unsigned long long* cand_val = nullptr;
  int64_t* cand_idx = nullptr;
  if (argmax_out && fixed) {
    IMPB_CUDA(cudaMallocAsync(&cand_val, sizeof(unsigned long long) * batch * cand_tiles, st));
    IMPB_CUDA(cudaMallocAsync(&cand_idx, sizeof(int64_t) * batch * cand_tiles, st));
  }
  int nblk = 0;
  if (blockmax_out) {
    if (!fixed) { set_error("impb_deconv_execute: this plan has no block maxima (impb_deconv_blockmax_info)"); return IMPB_EUNSUPPORTED; }
    nblk = (pl->precision == 64 ? pl->e64.fft.p : pl->e32.fft.p).col.radix[0];  // R0 of the column pass
    IMPB_CUDA(cudaMemsetAsync(blockmax_out, 0, sizeof(double) * batch * nblk, st));
  }
  if (pl->precision == 64 && !out_f64) {  // the FP64 engine stores doubles: a float32 `out` would be overrun 2x
    set_error("impb_deconv_execute: a precision-64 plan writes float64 output (out_f64 must be 1)");
    return IMPB_EINVAL;
  }
  const int64_t pairs = (batch + 1) / 2;
  const int ns = pl->nstreams;
  const cudaStream_t user = st;
  if (ns > 1) {
    IMPB_CUDA(cudaEventRecord(pl->fork_ev, user));
    for (int i = 0; i < ns; ++i) IMPB_CUDA(cudaStreamWaitEvent(pl->streams[i], pl->fork_ev, 0));
  }
  const int64_t slice_pairs = std::min<int64_t>(pairs, pl->chunk_pairs);
  int64_t chunk_idx = 0;
  for (int64_t p0 = 0; p0 < pairs; p0 += pl->chunk_pairs, ++chunk_idx) {
    const int np = (int)std::min<int64_t>(pl->chunk_pairs, pairs - p0);
    const int si = ns > 1 ? (int)(chunk_idx % ns) : 0;
    if (ns > 1) st = pl->streams[si];
    int rc;
    if (pl->precision == 64) {
      DeconvArgs<double> a{};
      const Plan2D& p = pl->e64.fft.p;
      a.ws = (double2*)workspace + (size_t)si * slice_pairs * p.N; a.tw = pl->e64.fft.tw; a.hs = pl->e64.hs; a.dcol = pl->e64.fft.dcol; a.drow = pl->e64.fft.drow;
      a.N = p.N; a.N1 = p.N1; a.N2 = p.N2;
      a.rec = rec; a.in_offsets = in_offsets; a.ld_in = ld_in; a.len = len; a.nrows = batch; a.row0 = 2 * p0; a.in_fmt = in_fmt;
      a.out = out; a.ld_out = ld_out; a.out_start = pl->s + out_offset; a.out_len = out_len; a.out_f64 = 1;
      a.absmax = (unsigned long long*)absmax_out;
      a.cand_val = cand_val; a.cand_idx = cand_idx; a.cand_tiles = cand_tiles;
      a.blockmax = (unsigned long long*)blockmax_out; a.nblk = nblk;
      rc = pl->e64.launch(a, np, st);
    } else {
      DeconvArgs<float> a{};
      const Plan2D& p = pl->e32.fft.p;
      a.ws = (float2*)workspace + (size_t)si * slice_pairs * p.N; a.tw = pl->e32.fft.tw; a.hs = pl->e32.hs; a.dcol = pl->e32.fft.dcol; a.drow = pl->e32.fft.drow;
      a.N = p.N; a.N1 = p.N1; a.N2 = p.N2;
      a.rec = rec; a.in_offsets = in_offsets; a.ld_in = ld_in; a.len = len; a.nrows = batch; a.row0 = 2 * p0; a.in_fmt = in_fmt;
      a.out = out; a.ld_out = ld_out; a.out_start = pl->s + out_offset; a.out_len = out_len; a.out_f64 = out_f64 ? 1 : 0;
      a.absmax = (unsigned long long*)absmax_out;
      a.cand_val = cand_val; a.cand_idx = cand_idx; a.cand_tiles = cand_tiles;
      a.blockmax = (unsigned long long*)blockmax_out; a.nblk = nblk;
      rc = pl->e32.launch(a, np, st);
    }
    if (rc) return rc;
  }
  if (ns > 1) {
    for (int i = 0; i < ns; ++i) {
      IMPB_CUDA(cudaEventRecord(pl->join_ev[i], pl->streams[i]));
      IMPB_CUDA(cudaStreamWaitEvent(user, pl->join_ev[i], 0));
    }
  }
  if (argmax_out) {
    if (fixed) {
      k_argmax_pick<<<(unsigned)((batch + 7) / 8), 256, 0, user>>>((const unsigned long long*)absmax_out, cand_val,
                                                                       cand_idx, cand_tiles, argmax_out, batch);
      IMPB_CUDA(cudaGetLastError());
      IMPB_CUDA(cudaFreeAsync(cand_val, user));
      IMPB_CUDA(cudaFreeAsync(cand_idx, user));
    } else {
      IMPB_CUDA(cudaMemsetAsync(argmax_out, 0xff, sizeof(int64_t) * batch, user));
      const dim3 grid((unsigned)batch, (unsigned)std::max<int64_t>(1, std::min<int64_t>(out_len / 4096, 1024)));
      if (pl->precision == 64 || out_f64)
        k_row_argmax<double><<<grid, 256, 0, user>>>((const double*)out, ld_out, out_len,
                                                     (const unsigned long long*)absmax_out, (unsigned long long*)argmax_out);
      else
        k_row_argmax<float><<<grid, 256, 0, user>>>((const float*)out, ld_out, out_len,
                                                    (const unsigned long long*)absmax_out, (unsigned long long*)argmax_out);
      IMPB_CUDA(cudaGetLastError());
    }
  }
  return IMPB_OK;
}

}  // extern "C"
