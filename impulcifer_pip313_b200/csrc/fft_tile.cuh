// Tile FFT engine for sm_100a: mixed-radix Stockham-free (in-place DIF / DIT) sub-FFTs on a shared-memory
// tile, the building block of every large transform in this library.
//
// A length-N transform is factored N = N1 x N2 ("four-step").  One CTA owns a tile of TILE adjacent
// columns (column pass, sub-FFT length N1, element stride N2 in HBM) or TILE adjacent rows (row pass,
// sub-FFT length N2, contiguous in HBM).  TILE*sizeof(complex) = 128 B, so every HBM access of a warp is
// made of whole 128-byte segments.  The tile lives in shared memory XOR-swizzled ([S][TILE] for the column
// pass, [TILE][S] for the row pass, see tile_addr) so that a quarter-warp hits 8 distinct 16-byte bank groups.
//
// Forward sub-FFTs are decimation-in-frequency (natural order in, digit-reversed order out), inverse
// sub-FFTs are the exact mirror decimation-in-time (digit-reversed in, natural out).  Convolutions and the
// cepstral pipeline only ever do elementwise work between a forward and an inverse transform, so the data
// is never reordered: spectra stay in "scrambled" position order and the host permutes the constant
// tables (inverse-sweep spectrum, twiddles) instead.
//
// The first stage reads straight from HBM into registers and the last stage writes straight back, so a
// pass with m radix stages touches shared memory 2(m-1) times per element.  Butterflies of radix up to 16
// run in registers with compile-time twiddles (twiddle_consts.cuh); inter-stage twiddles come from a
// quarter-wave table kept in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "twiddle_consts.cuh"

namespace impb {

constexpr int MAX_STAGES = 6;

template <typename T> struct CxOf;
template <> struct CxOf<double> { using type = double2; };
template <> struct CxOf<float> { using type = float2; };
template <typename T> using cx_t = typename CxOf<T>::type;

template <typename T> struct TileOf { static constexpr int value = 64 / sizeof(T); };  // 8 (f64) / 16 (f32)

template <typename T> __host__ __device__ __forceinline__ cx_t<T> mkc(T a, T b) {
  cx_t<T> r; r.x = a; r.y = b; return r;
}
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
  C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// FP32: Blackwell's packed FP32x2 instructions (FADD2 / FFMA2 / FMUL2, sm_100+) do the re and im halves of a complex
// add / subtract in ONE issue slot; the FP32 K1 kernels are issue-bound with FADD as their most frequent instruction
// (profiles/r02_k1_ncu_full_f32.md).  Same IEEE results as the scalar forms (a - b as fma(b, -1, a) is exact-then-rounded).
#ifndef IMPB_NO_PACKED_F32
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
#ifdef IMPB_PACKED_F32_CMUL
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return __ffma2_rn(make_float2(a.x, a.x), b, __fmul2_rn(make_float2(a.y, a.y), make_float2(-b.y, b.x)));
}
#endif
#endif
template <typename C> __device__ __forceinline__ C cconj(C a) { a.y = -a.y; return a; }
// multiply by -i (forward quarter turn) / +i
template <typename C> __device__ __forceinline__ C mul_mi(C a) { C r; r.x = a.y; r.y = -a.x; return r; }
template <typename C> __device__ __forceinline__ C mul_pi(C a) { C r; r.x = -a.y; r.y = a.x; return r; }
template <bool INV, typename C> __device__ __forceinline__ C rot90(C a) { return INV ? mul_pi(a) : mul_mi(a); }

// ------------------------------------------------------------------------------------------------
// in-register DFTs, natural order in and out.  INV selects exp(+i..) (unnormalised inverse).
// ------------------------------------------------------------------------------------------------
template <typename T, int R, bool INV> struct Dft;

template <typename T, bool INV> struct Dft<T, 1, INV> {
  __device__ __forceinline__ static void run(cx_t<T>*) {}
};

template <typename T, bool INV> struct Dft<T, 2, INV> {
  __device__ __forceinline__ static void run(cx_t<T>* x) {
    cx_t<T> a = x[0], b = x[1];
    x[0] = cadd(a, b);
    x[1] = csub(a, b);
  }
};

template <typename T, bool INV> struct Dft<T, 3, INV> {
  __device__ __forceinline__ static void run(cx_t<T>* x) {
    const T k = (T)0.8660254037844386467637231707529362;
    cx_t<T> t = cadd(x[1], x[2]);
    cx_t<T> d = csub(x[1], x[2]);
    cx_t<T> m = mkc<T>(x[0].x - (T)0.5 * t.x, x[0].y - (T)0.5 * t.y);
    cx_t<T> r = rot90<INV>(mkc<T>(k * d.x, k * d.y));
    x[0] = cadd(x[0], t);
    x[1] = cadd(m, r);
    x[2] = csub(m, r);
  }
};

template <typename T, bool INV> struct Dft<T, 4, INV> {
  __device__ __forceinline__ static void run(cx_t<T>* x) {
    cx_t<T> s02 = cadd(x[0], x[2]), d02 = csub(x[0], x[2]);
    cx_t<T> s13 = cadd(x[1], x[3]), d13 = rot90<INV>(csub(x[1], x[3]));
    x[0] = cadd(s02, s13);
    x[2] = csub(s02, s13);
    x[1] = cadd(d02, d13);
    x[3] = csub(d02, d13);
  }
};

template <typename T, bool INV> struct Dft<T, 5, INV> {
  __device__ __forceinline__ static void run(cx_t<T>* x) {
    const T c1 = (T)0.3090169943749474241022934171828191;
    const T c2 = (T)-0.8090169943749474241022934171828191;
    const T s1 = (T)0.9510565162951535721164393333793821;
    const T s2 = (T)0.5877852522924731291687059546390728;
    cx_t<T> t1 = cadd(x[1], x[4]), t3 = csub(x[1], x[4]);
    cx_t<T> t2 = cadd(x[2], x[3]), t4 = csub(x[2], x[3]);
    cx_t<T> m1 = mkc<T>(x[0].x + c1 * t1.x + c2 * t2.x, x[0].y + c1 * t1.y + c2 * t2.y);
    cx_t<T> m2 = mkc<T>(x[0].x + c2 * t1.x + c1 * t2.x, x[0].y + c2 * t1.y + c1 * t2.y);
    cx_t<T> n1 = rot90<INV>(mkc<T>(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
    cx_t<T> n2 = rot90<INV>(mkc<T>(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
    x[0] = cadd(x[0], cadd(t1, t2));
    x[1] = cadd(m1, n1);
    x[4] = csub(m1, n1);
    x[2] = cadd(m2, n2);
    x[3] = csub(m2, n2);
  }
};

// multiply by the compile-time constant exp(-/+ 2 pi i m / R); trivial rotations cost nothing
template <typename T, int R, bool INV> __device__ __forceinline__ cx_t<T> mul_wconst(cx_t<T> z, int m) {
  const double c = TwK<R>::c(m), s = TwK<R>::s(m);
  if (s == 0.0 && c == 1.0) return z;
  if (s == 0.0 && c == -1.0) return mkc<T>(-z.x, -z.y);
  if (c == 0.0 && s == 1.0) return rot90<INV>(z);
  if (c == 0.0 && s == -1.0) return rot90<!INV>(z);
  const T wc = (T)c, ws = INV ? (T)s : (T)(-s);
  return mkc<T>(z.x * wc - z.y * ws, z.x * ws + z.y * wc);
}

// R = A*B Cooley-Tukey in registers:  n = B a + b,  k = ka + A kb
template <typename T, int A, int B, bool INV> __device__ __forceinline__ void dft_ab(cx_t<T>* x) {
  constexpr int R = A * B;
  cx_t<T> t[R];
#pragma unroll
  for (int b = 0; b < B; ++b) {
    cx_t<T> u[A];
#pragma unroll
    for (int a = 0; a < A; ++a) u[a] = x[B * a + b];
    Dft<T, A, INV>::run(u);
#pragma unroll
    for (int ka = 0; ka < A; ++ka) t[b * A + ka] = mul_wconst<T, R, INV>(u[ka], (b * ka) % R);
  }
#pragma unroll
  for (int ka = 0; ka < A; ++ka) {
    cx_t<T> v[B];
#pragma unroll
    for (int b = 0; b < B; ++b) v[b] = t[b * A + ka];
    Dft<T, B, INV>::run(v);
#pragma unroll
    for (int kb = 0; kb < B; ++kb) x[ka + A * kb] = v[kb];
  }
}

#define IMPB_DFT_AB(R_, A_, B_)                                                       \
  template <typename T, bool INV> struct Dft<T, R_, INV> {                            \
    __device__ __forceinline__ static void run(cx_t<T>* x) { dft_ab<T, A_, B_, INV>(x); } \
  };
IMPB_DFT_AB(6, 2, 3)
IMPB_DFT_AB(8, 2, 4)
IMPB_DFT_AB(9, 3, 3)
IMPB_DFT_AB(10, 2, 5)
IMPB_DFT_AB(12, 3, 4)
IMPB_DFT_AB(15, 3, 5)
IMPB_DFT_AB(16, 4, 4)
#undef IMPB_DFT_AB

// ------------------------------------------------------------------------------------------------
// sub-FFT description (host fills it, kernels take it by value)
// ------------------------------------------------------------------------------------------------
struct SubPlan {
  int S;                   // sub-FFT length
  int nst;                 // number of radix stages
  int qmode;               // 1: quarter-wave twiddle table (S % 4 == 0), 0: full table
  int ntab;                // entries in the table (S/4 or S)
  int radix[MAX_STAGES];
  int len[MAX_STAGES];     // block length entering stage s (len[0] = S)
  int stride[MAX_STAGES];  // len / radix
  int mult[MAX_STAGES];    // S / len (twiddle index multiplier)
};

// exp(-2 pi i idx / S) from the shared-memory table; idx in [0, S)
template <typename T> __device__ __forceinline__ cx_t<T> tw_fetch(const cx_t<T>* tab, int idx, const SubPlan& sp) {
  if (!sp.qmode) return tab[idx];
  const int q4 = sp.ntab;
  const int quad = (idx >= q4) + (idx >= 2 * q4) + (idx >= 3 * q4);
  const cx_t<T> v = tab[idx - quad * q4];  // (cos phi, sin phi)
  switch (quad) {
    case 0: return mkc<T>(v.x, -v.y);
    case 1: return mkc<T>(-v.y, -v.x);
    case 2: return mkc<T>(-v.x, v.y);
    default: return mkc<T>(v.y, v.x);
  }
}

// ------------------------------------------------------------------------------------------------
// work-item order and shared-memory layout of a tile
//  * column pass (ROW = false): the TILE columns are the fast index of a work item, the tile is [S][TILE] with
//    slot = c ^ (e & (TILE-1)): a quarter-warp is one row x 8 columns = one 128-byte HBM segment and 8 bank groups.
//  * row pass (ROW = true): the butterfly index j is the fast index, the tile is [TILE][SP] (SP = S rounded up
//    to 8) with e' = e ^ ((e >> 3) & 7): a quarter-warp is 8 consecutive elements of ONE row, again one 128-byte
//    segment (with the columns fast it would be 8 different lines per request).
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int tile_row_pitch(int S) { return (S + 7) & ~7; }

template <bool ROW, int TILE> __device__ __forceinline__ void tile_decode(int w, int per_col, int& c, int& j) {
  if (ROW) { c = w / per_col; j = w - c * per_col; } else { c = w % TILE; j = w / TILE; }
}
template <bool ROW, int TILE> __device__ __forceinline__ int tile_addr(int e, int c, int SP) {
  return ROW ? c * SP + (e ^ ((e >> 3) & 7)) : e * TILE + (c ^ (e & (TILE - 1)));
}

// all items of one radix-R stage.  ld(e, c) / st(e, c, v) hide whether the operand is HBM (first / last stage) or
// the shared tile.  Stage twiddles w_p = exp(-2 pi i np p / len) come from the quarter-wave table in shared memory.
template <typename T, int TILE, int R, bool INV, bool ROW, class LD, class ST>
__device__ __forceinline__ void stage_loop(const SubPlan& sp, int s, const cx_t<T>* tab, LD ld, ST st) {
  const int stride = sp.stride[s], len = sp.len[s], mult = sp.mult[s];
  const int per_col = sp.S / R;
  const int items = per_col * TILE;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    int c, j;
    tile_decode<ROW, TILE>(w, per_col, c, j);
    const int b = j / stride;
    const int np = j - b * stride;
    const int base = b * len + np;
    cx_t<T> x[R];
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = ld(base + q * stride, c);
    if (stride > 1) {
      // every twiddle is a correctly rounded table entry: the cepstral (minimum-phase) pipeline takes logs of
      // spectra near their nulls and amplifies each extra ulp of transform noise (min-phase tolerance)
      cx_t<T> wv[R];
      const int step = np * mult;
#pragma unroll
      for (int p = 1; p < R; ++p) wv[p] = tw_fetch<T>(tab, step * p, sp);
      if (INV) {
#pragma unroll
        for (int p = 1; p < R; ++p) x[p] = cmul(x[p], cconj(wv[p]));
        Dft<T, R, true>::run(x);
      } else {
        Dft<T, R, false>::run(x);
#pragma unroll
        for (int p = 1; p < R; ++p) x[p] = cmul(x[p], wv[p]);
      }
    } else {
      Dft<T, R, INV>::run(x);
    }
#pragma unroll
    for (int p = 0; p < R; ++p) st(base + p * stride, c, x[p]);
  }
}

#define IMPB_RADIX_SWITCH(r_, CALL)   \
  switch (r_) {                       \
    case 2: { CALL(2); } break;       \
    case 3: { CALL(3); } break;       \
    case 4: { CALL(4); } break;       \
    case 5: { CALL(5); } break;       \
    case 6: { CALL(6); } break;       \
    case 8: { CALL(8); } break;       \
    case 9: { CALL(9); } break;       \
    case 10: { CALL(10); } break;     \
    case 12: { CALL(12); } break;     \
    case 15: { CALL(15); } break;     \
    case 16: { CALL(16); } break;     \
    default: break;                   \
  }

template <typename T> __device__ __forceinline__ void load_table(cx_t<T>* tab_s, const cx_t<T>* tab_g, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) tab_s[i] = tab_g[i];
}

// ------------------------------------------------------------------------------------------------
// forward (DIF) tile transform:  gl(e, c) -> [stages] -> gs(pos, c, v)
// ------------------------------------------------------------------------------------------------
template <typename T, int TILE, bool ROW, class GL, class GS>
__device__ __forceinline__ void tile_fft_forward(const SubPlan& sp, cx_t<T>* tile, const cx_t<T>* tab, GL gl, GS gs) {
  const int SP = tile_row_pitch(sp.S);
  for (int s = 0; s < sp.nst; ++s) {
    const bool first = (s == 0), last = (s == sp.nst - 1);
    auto ld = [&](int e, int c) -> cx_t<T> { return first ? gl(e, c) : tile[tile_addr<ROW, TILE>(e, c, SP)]; };
    auto st = [&](int e, int c, cx_t<T> v) {
      if (last) gs(e, c, v); else tile[tile_addr<ROW, TILE>(e, c, SP)] = v;
    };
#define IMPB_CALL(R_) stage_loop<T, TILE, R_, false, ROW>(sp, s, tab, ld, st)
    IMPB_RADIX_SWITCH(sp.radix[s], IMPB_CALL)
#undef IMPB_CALL
    if (!last) __syncthreads();
  }
}

// inverse (DIT) tile transform, the mirror image: stages run nst-1 .. 0
template <typename T, int TILE, bool ROW, class GL, class GS>
__device__ __forceinline__ void tile_fft_inverse(const SubPlan& sp, cx_t<T>* tile, const cx_t<T>* tab, GL gl, GS gs) {
  const int SP = tile_row_pitch(sp.S);
  for (int s = sp.nst - 1; s >= 0; --s) {
    const bool first = (s == sp.nst - 1), last = (s == 0);
    auto ld = [&](int e, int c) -> cx_t<T> { return first ? gl(e, c) : tile[tile_addr<ROW, TILE>(e, c, SP)]; };
    auto st = [&](int e, int c, cx_t<T> v) {
      if (last) gs(e, c, v); else tile[tile_addr<ROW, TILE>(e, c, SP)] = v;
    };
#define IMPB_CALL(R_) stage_loop<T, TILE, R_, true, ROW>(sp, s, tab, ld, st)
    IMPB_RADIX_SWITCH(sp.radix[s], IMPB_CALL)
#undef IMPB_CALL
    if (!last) __syncthreads();
  }
}

// The last DIF stage, a pointwise multiply and the first DIT stage touch the same r elements from the
// same thread: done in registers with no shared-memory round trip in between.  Row pass only.
template <typename T, int TILE, int R, class LD, class ST, class PW>
__device__ __forceinline__ void mid_loop(const SubPlan& sp, LD ld, ST st, PW pw) {
  const int per_col = sp.S / R;
  const int items = per_col * TILE;
  for (int w = threadIdx.x; w < items; w += blockDim.x) {
    int c, j;
    tile_decode<true, TILE>(w, per_col, c, j);
    cx_t<T> x[R];
    const int base = j * R;
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = ld(base + q, c);
    Dft<T, R, false>::run(x);
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = cmul(x[q], pw(q, w));
    Dft<T, R, true>::run(x);
#pragma unroll
    for (int q = 0; q < R; ++q) st(base + q, c, x[q]);
  }
}

// forward -> pointwise multiply -> inverse on one tile (circular convolution along the tile's rows)
template <typename T, int TILE, class GL, class GS, class PW>
__device__ __forceinline__ void tile_fft_conv(const SubPlan& sp, cx_t<T>* tile, const cx_t<T>* tab, GL gl, GS gs, PW pw) {
  const int SP = tile_row_pitch(sp.S);
  const int m = sp.nst - 1;
  auto sl = [&](int e, int c) -> cx_t<T> { return tile[tile_addr<true, TILE>(e, c, SP)]; };
  auto ss = [&](int e, int c, cx_t<T> v) { tile[tile_addr<true, TILE>(e, c, SP)] = v; };
  for (int s = 0; s < m; ++s) {
    const bool first = (s == 0);
    auto ld = [&](int e, int c) -> cx_t<T> { return first ? gl(e, c) : sl(e, c); };
#define IMPB_CALL(R_) stage_loop<T, TILE, R_, false, true>(sp, s, tab, ld, ss)
    IMPB_RADIX_SWITCH(sp.radix[s], IMPB_CALL)
#undef IMPB_CALL
    __syncthreads();
  }
  {
    const bool only = (m == 0);
    auto ld = [&](int e, int c) -> cx_t<T> { return only ? gl(e, c) : sl(e, c); };
    auto st = [&](int e, int c, cx_t<T> v) {
      if (only) gs(e, c, v); else ss(e, c, v);
    };
#define IMPB_CALL(R_) mid_loop<T, TILE, R_>(sp, ld, st, pw)
    IMPB_RADIX_SWITCH(sp.radix[m], IMPB_CALL)
#undef IMPB_CALL
    if (!only) __syncthreads();
  }
  for (int s = m - 1; s >= 0; --s) {
    const bool last = (s == 0);
    auto st = [&](int e, int c, cx_t<T> v) {
      if (last) gs(e, c, v); else ss(e, c, v);
    };
#define IMPB_CALL(R_) stage_loop<T, TILE, R_, true, true>(sp, s, tab, sl, st)
    IMPB_RADIX_SWITCH(sp.radix[s], IMPB_CALL)
#undef IMPB_CALL
    if (!last) __syncthreads();
  }
}

template <typename T> inline size_t tile_smem_bytes(const SubPlan& sp, int tile) {
  return ((size_t)tile_row_pitch(sp.S) * tile + (size_t)sp.ntab) * sizeof(cx_t<T>);
}

}  // namespace impb
