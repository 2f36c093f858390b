// Compile-time specialised tile FFTs: same algorithm and memory layout as fft_tile.cuh, but the radix
// sequence <R0,R1,R2> is a template parameter, so every stride, shared-memory offset and loop trip count
// is an immediate.  Inter-stage twiddles come from small per-stage tables tw_s[np][p-1] (read through the
// read-only path; the 8 columns of a tile share each entry, so a warp touches one or two sectors).
// Used for the sub-FFT lengths the headline workloads need; every other length takes the generic path.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "fft_tile.cuh"

#ifndef IMPB_TRIP_UNROLL
#define IMPB_TRIP_UNROLL 1
#endif

namespace impb {

constexpr int kTripUnroll = IMPB_TRIP_UNROLL;
#ifndef IMPB_STAGE_UNROLL
#define IMPB_STAGE_UNROLL 4  // copies of the cp.async staging loop of pass A issued back to back
#endif
constexpr int kStageUnroll = IMPB_STAGE_UNROLL;
#ifndef IMPB_PREFETCH
#define IMPB_PREFETCH 0
#endif
constexpr bool kPrefetch = IMPB_PREFETCH != 0;

template <int R0_, int R1_, int R2_> struct Radix3 {
  static constexpr int R0 = R0_, R1 = R1_, R2 = R2_;
  static constexpr int S = R0_ * R1_ * R2_;
  static constexpr int S0 = S / R0_;   // stride of stage 0 == length entering stage 1
  static constexpr int S1 = S0 / R1_;  // stride of stage 1 == R2
  static constexpr int TW0 = S0 * (R0_ - 1);  // entries of the stage-0 table
  static constexpr int TW1 = S1 * (R1_ - 1);
};

// Work-item -> (column c, butterfly j) decode and shared-memory addressing.
//  * column pass (ROW = false): c fastest.  A quarter-warp (8 lanes x 16 B) is one row x 8 columns: one 128-byte
//    HBM segment, and 8 distinct bank groups through the XOR swizzle  slot = c ^ (e & 7).
//  * row pass (ROW = true): j fastest.  A quarter-warp is 8 consecutive elements of ONE row: again one 128-byte
//    HBM segment (with c fastest it would be 8 different lines = 8x the LSU wavefronts, measured 18 wavefronts
//    per request in round 1).  Tile is [c][S] with  e' = e ^ ((e >> 3) & 7), conflict-free for strides 64, 8, 1.
// The row pitch is S rounded up to 16 (the XOR stays inside an aligned group of 8 / 16 elements), so sub-FFT lengths
// that are not multiples of 16 (125, 75 ...) do not run into the next row; for 768 / 256 / 128 there is no padding.
constexpr int fx_pitch(int S) { return (S + 15) & ~15; }
template <int S, int TILE> constexpr size_t fx_tile_elems() { return (size_t)fx_pitch(S) * TILE; }
template <bool ROW, int TILE, int PER_COL> __device__ __forceinline__ void fx_decode(int w, int& c, int& j) {
  if (ROW) { j = w % PER_COL; c = w / PER_COL; } else { c = w % TILE; j = w / TILE; }
}
template <bool ROW, int TILE, int S> __device__ __forceinline__ int fx_smem(int e, int c) {
  if (!ROW) return e * TILE + (c ^ (e & (TILE - 1)));
  constexpr int SP = fx_pitch(S);
  if (TILE == 8) return c * SP + (e ^ ((e >> 3) & 7));  // 16-byte elements: 8 slots per 128 B, quarter-warps
  // 8-byte elements (FP32 tiles): a half-warp must cover 16 distinct slots for strides 96, 8 and 1
  const int v = e >> 4;
  return c * SP + (e ^ (v & 7) ^ (((v >> 1) & 1) << 3));
}

// Stage twiddles w_p = exp(-2 pi i np p / LEN), p = 1..R-1, from the power-of-two entries of the table only
// (w_1, w_2, w_4, w_8: 3-4 loads instead of R-1; the rest are single products of table values, <= 2 roundings).
// Table layout: entry k of butterfly np holds w_(2^k):  column pass tw[np][k], row pass tw[k][np] (coalesced).
#ifndef IMPB_TW_LOADS
#define IMPB_TW_LOADS 1  // how many of the power-of-two entries are loaded; the rest are squared from the last one
#endif
// FP32 tiles are issue-bound, not load-bound (profiles/r02_k1_ncu_full_f32.md: complex multiplies are a third of the
// issued instructions of the row pass), so they load all the power-of-two entries and compose only the rest:
// K1 FP32 6.92 -> 6.73 ms per 1792 IRs; the same switch costs FP64 (load-bound) 2 %.
#ifndef IMPB_TW_LOADS_F32
#define IMPB_TW_LOADS_F32 4
#endif
// FULL: every w_p is its own correctly rounded table entry (tw[np][p-1] / tw[p-1][np]).  Used by the frequency-
// response transforms: the cepstral minimum-phase chain takes logs of spectra next to their nulls and amplifies
// each extra ulp of transform noise (min-phase tolerance), so no composed twiddles there.
template <typename T, int R, bool ROW, int STRIDE, bool FULL = false>
__device__ __forceinline__ void fx_twiddles(const cx_t<T>* __restrict__ tw, int np, cx_t<T>* w) {
  if constexpr (FULL) {
#pragma unroll
    for (int p = 1; p < R; ++p) w[p] = __ldg(ROW ? tw + (p - 1) * STRIDE + np : tw + np * (R - 1) + (p - 1));
    return;
  }
  constexpr int NB = (R > 8) ? 4 : (R > 4) ? 3 : (R > 2) ? 2 : 1;
  constexpr int WANT = sizeof(T) == 4 ? IMPB_TW_LOADS_F32 : IMPB_TW_LOADS;
  constexpr int NL = NB < WANT ? NB : WANT;
  const cx_t<T>* twp = ROW ? tw + np : tw + np * NB;
  constexpr int TWS = ROW ? STRIDE : 1;
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    if ((1 << k) >= R) continue;
    if (k < NL) w[1 << k] = __ldg(twp + k * TWS);
    else w[1 << k] = cmul(w[1 << (k - 1)], w[1 << (k - 1)]);
  }
#pragma unroll
  for (int p = 3; p < R; ++p) {
    const int hi = (p >= 8) ? 8 : (p >= 4) ? 4 : 2;
    if (p != hi) w[p] = cmul(w[hi], w[p - hi]);
  }
}

// one radix-R stage over the whole tile; LD/ST address elements as (e, c).  POST (optional) supplies one
// multiplier per OUTPUT element (four-step twiddle); it is fetched at the top of the trip so that its latency
// hides behind the butterfly instead of sitting in front of the store.
struct NoPost {};
// A POST functor with `static constexpr bool progression = true` supplies first(e, c) = the multiplier of a butterfly's
// output 0 and step(c) = D with multiplier(output p) = first * D^p (four-step twiddles: the R outputs of the stage sit
// on an arithmetic progression of frequencies / samples): ONE table load per butterfly instead of R.
template <class P, class = void> struct is_progression : std::false_type {};
template <class P> struct is_progression<P, std::enable_if_t<P::progression>> : std::true_type {};
struct NoTripHook { __device__ __forceinline__ void operator()(int) const {} };
// TRIP (optional) is called with the trip number before the trip's loads: the TMA-fed column pass waits there for the
// box of rows the trip reads, so the butterflies of the first rows overlap the transfer of the rest of the tile.
template <typename T, int TILE, int NT, int S, int LEN, int R, bool INV, bool ROW, bool PF = false, bool TWFULL = false,
          class LD, class ST, class POST, class TRIP = NoTripHook>
__device__ __forceinline__ void fx_stage(const cx_t<T>* __restrict__ tw, LD ld, ST st, POST post, TRIP trip = TRIP{}) {
  constexpr int STRIDE = LEN / R;
  constexpr int PER_COL = S / R;
  constexpr int ITEMS = PER_COL * TILE;
  constexpr int TRIPS = (ITEMS + NT - 1) / NT;
  constexpr bool HAS_POST = !std::is_same<POST, NoPost>::value;
  constexpr bool PROG = is_progression<POST>::value;
  constexpr int NPOW = (R > 8) ? 4 : (R > 4) ? 3 : (R > 2) ? 2 : 1;
  cx_t<T> dpow[PROG ? NPOW : 1];
  if constexpr (PROG && !ROW) {
    // column pass: a thread keeps its column in every trip (NT is a multiple of TILE), so D is fetched once
    static_assert(NT % TILE == 0, "per-thread constant column");
    dpow[0] = post.step(threadIdx.x % TILE);
#pragma unroll
    for (int k = 1; k < NPOW; ++k) dpow[k] = cmul(dpow[k - 1], dpow[k - 1]);
  }
  // PF (stages that read HBM): the operands of trip it+1 are requested before trip it is computed
  cx_t<T> xn[PF ? R : 1];
  if constexpr (PF) {
    int c, j;
    fx_decode<ROW, TILE, PER_COL>(threadIdx.x, c, j);
    const int base = (j / STRIDE) * LEN + (j % STRIDE);
    if ((ITEMS % NT) == 0 || threadIdx.x < ITEMS) {
#pragma unroll
      for (int q = 0; q < R; ++q) xn[q] = ld(base + q * STRIDE, c);
    }
  }
#pragma unroll kTripUnroll
  for (int it = 0; it < TRIPS; ++it) {
    trip(it);
    const int w = threadIdx.x + it * NT;
    if ((ITEMS % NT) != 0 && w >= ITEMS) break;
    int c, j;
    fx_decode<ROW, TILE, PER_COL>(w, c, j);
    const int b = j / STRIDE, np = j % STRIDE;
    const int base = b * LEN + np;
    cx_t<T> m[(HAS_POST && !PROG) ? R : 1];
    cx_t<T> m0;  // PROG: multiplier of output 0; output p takes m0 * D^p (D, D^2, D^4 ... fetched once per thread)
    if constexpr (HAS_POST && !PROG) {
#pragma unroll
      for (int p = 0; p < R; ++p) m[p] = post(base + p * STRIDE, c);
    }
    if constexpr (PROG) {
      m0 = post.first(base, c);
      if constexpr (ROW) {  // row pass: the row changes from trip to trip, D with it (one more load per butterfly)
        dpow[0] = post.step(c);
#pragma unroll
        for (int k = 1; k < NPOW; ++k) dpow[k] = cmul(dpow[k - 1], dpow[k - 1]);
      }
    }
    cx_t<T> x[R];
    if constexpr (PF) {
#pragma unroll
      for (int q = 0; q < R; ++q) x[q] = xn[q];
      const int w2 = w + NT;
      if (it + 1 < TRIPS && ((ITEMS % NT) == 0 || w2 < ITEMS)) {
        int c2, j2;
        fx_decode<ROW, TILE, PER_COL>(w2, c2, j2);
        const int base2 = (j2 / STRIDE) * LEN + (j2 % STRIDE);
#pragma unroll
        for (int q = 0; q < R; ++q) xn[q] = ld(base2 + q * STRIDE, c2);
      }
    } else {
#pragma unroll
      for (int q = 0; q < R; ++q) x[q] = ld(base + q * STRIDE, c);
    }
#ifdef IMPB_ABL_NO_MATH  // developer ablation: loads, exchanges and stores only
    if (false) {
#else
    if (STRIDE > 1) {
#endif
      cx_t<T> wv[R];
      fx_twiddles<T, R, ROW, STRIDE, TWFULL>(tw, np, wv);
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
#ifndef IMPB_ABL_NO_MATH
      Dft<T, R, INV>::run(x);
#endif
    }
    if constexpr (HAS_POST && !PROG) {
#pragma unroll
      for (int p = 0; p < R; ++p) x[p] = cmul(x[p], m[p]);
    }
    if constexpr (PROG) {
      // m_p = m_(p - hi) * D^hi with hi the largest power of two <= p: every multiplier is a product of at most
      // log2(R) table-derived factors (the same composition the stage twiddles use)
      cx_t<T> mp[R];
      mp[0] = m0;
#pragma unroll
      for (int p = 1; p < R; ++p) {
        const int hi = (p >= 8) ? 8 : (p >= 4) ? 4 : (p >= 2) ? 2 : 1;
        const int k = (hi == 8) ? 3 : (hi == 4) ? 2 : (hi == 2) ? 1 : 0;
        mp[p] = cmul(mp[p - hi], dpow[k]);
      }
#pragma unroll
      for (int p = 0; p < R; ++p) x[p] = cmul(x[p], mp[p]);
    }
#pragma unroll
    for (int p = 0; p < R; ++p) {
      // a store functor may ask for the (compile-time, after unrolling) butterfly output number p
      if constexpr (std::is_invocable_v<ST, int, int, cx_t<T>, int>) st(base + p * STRIDE, c, x[p], p);
      else st(base + p * STRIDE, c, x[p]);
    }
  }
}

// last forward stage x pointwise x first inverse stage, in registers (row pass only)
template <typename T, int TILE, int NT, int S, int R, class LD, class ST, class PW>
__device__ __forceinline__ void fx_mid(LD ld, ST st, PW pw) {
  constexpr int PER_COL = S / R;
  constexpr int ITEMS = PER_COL * TILE;
  constexpr int TRIPS = (ITEMS + NT - 1) / NT;
#pragma unroll kTripUnroll
  for (int it = 0; it < TRIPS; ++it) {
    const int w = threadIdx.x + it * NT;
    if ((ITEMS % NT) != 0 && w >= ITEMS) break;
    int c, j;
    fx_decode<true, TILE, PER_COL>(w, c, j);
    cx_t<T> h[R];
#pragma unroll
    for (int q = 0; q < R; ++q) h[q] = pw(q, w);
    cx_t<T> x[R];
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = ld(j * R + q, c);
#ifndef IMPB_ABL_NO_MATH
    Dft<T, R, false>::run(x);
#endif
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = cmul(x[q], h[q]);
#ifndef IMPB_ABL_NO_MATH
    Dft<T, R, true>::run(x);
#endif
#pragma unroll
    for (int q = 0; q < R; ++q) st(j * R + q, c, x[q]);
  }
}

// adapter: F::post(t, e, c) per output, or (when F offers post_first / post_step) the progression form
template <class F, class = void> struct has_prog : std::false_type {};
template <class F> struct has_prog<F, std::enable_if_t<F::post_progression>> : std::true_type {};
template <typename T, class F, bool = has_prog<F>::value> struct PostOf {
  const F& f;
  int t;
  __device__ __forceinline__ cx_t<T> operator()(int e, int c) const { return f.post(t, e, c); }
};
template <typename T, class F> struct PostOf<T, F, true> {
  static constexpr bool progression = true;
  const F& f;
  int t;
  __device__ __forceinline__ cx_t<T> operator()(int e, int c) const { return f.post(t, e, c); }
  __device__ __forceinline__ cx_t<T> first(int e, int c) const { return f.post(t, e, c); }
  __device__ __forceinline__ cx_t<T> step(int c) const { return f.post_step(t, c); }
};

template <typename T, int TILE, int NT, class RX, class F>
__global__ void __launch_bounds__(NT, 512 / NT)
k_fx_fwd(const cx_t<T>* __restrict__ tw0, const cx_t<T>* __restrict__ tw1, const F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx_t<T>* tile = reinterpret_cast<cx_t<T>*>(smem_raw);
  const int t = blockIdx.x, b = blockIdx.y;
  auto gl = [&](int e, int c) -> cx_t<T> { return f.load(b, t, e, c); };
  auto gs = [&](int e, int c, cx_t<T> v) { f.store(b, t, e, c, v); };
  auto sl = [&](int e, int c) -> cx_t<T> { return tile[fx_smem<false, TILE, RX::S>(e, c)]; };
  auto ss = [&](int e, int c, cx_t<T> v) { tile[fx_smem<false, TILE, RX::S>(e, c)] = v; };
  PostOf<T, F> post{f, t};
  fx_stage<T, TILE, NT, RX::S, RX::S, RX::R0, false, false>(tw0, gl, ss, NoPost{});
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S0, RX::R1, false, false>(tw1, sl, ss, NoPost{});
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S1, RX::R2, false, false>(nullptr, sl, gs, post);
}

// Column forward pass whose input tile is staged HBM -> shared memory with cp.async (LDGSTS): every load of
// the tile is in flight at once and none of them occupies a register, so the pass is bound by HBM bandwidth
// instead of (trips x HBM latency).  F::src(b, t, e, c, which) returns the address of the real sample that
// becomes the re (which = 0) / im (which = 1) part of element (e, c), or nullptr for a zero.
template <int BYTES> __device__ __forceinline__ void cp_async_n(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (BYTES == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}

template <typename T, int TILE, int NT, class RX, class F>
__global__ void __launch_bounds__(NT, 512 / NT)
k_fx_fwd_async(const cx_t<T>* __restrict__ tw0, const cx_t<T>* __restrict__ tw1, const F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx_t<T>* tile = reinterpret_cast<cx_t<T>*>(smem_raw);
  const int t = blockIdx.x, b = blockIdx.y;
  // adjacent lanes copy the re and the im part of one element: a warp instruction then fills 16 whole complex slots
  // (contiguous shared-memory bytes) instead of the halves of 32 (which cost twice the shared-memory wavefronts).
  // A thread keeps its column c and its re/im half for the whole tile and walks down the rows in steps of
  // (NT/2)/TILE, so the source address is ONE pointer advanced by a constant stride and the bounds test one
  // compare against a per-thread row limit: ~6 instructions per copy instead of the ~20 of a per-element
  // src(b, t, e, c, which) call (the loop used to issue as many instructions as the butterflies of the tile).
  {
    static_assert((NT / 2) % TILE == 0, "a thread must stay in one column");
    constexpr int ESTEP = (NT / 2) / TILE;
    const int which = threadIdx.x & 1, i0 = threadIdx.x >> 1;
    const int c = i0 % TILE, e0 = i0 / TILE;
    int e_lim;        // rows [0, e_lim) of this column hold samples, the rest is zero padding
    int64_t e_pitch;  // elements between consecutive rows of the column
    const T* col = f.src_column(b, t, c, which, &e_lim, &e_pitch);
    const T* src = col + (int64_t)e0 * e_pitch;
    const int64_t step = (int64_t)ESTEP * e_pitch;
#pragma unroll kStageUnroll
    for (int e = e0; e < RX::S; e += ESTEP, src += step) {
      T* dst = reinterpret_cast<T*>(&tile[fx_smem<false, TILE, RX::S>(e, c)]) + which;
      if (e < e_lim) cp_async_n<sizeof(T)>(dst, src); else *dst = (T)0;
    }
  }
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();
  auto gs = [&](int e, int c, cx_t<T> v) { f.store(b, t, e, c, v); };
  auto sl = [&](int e, int c) -> cx_t<T> { return tile[fx_smem<false, TILE, RX::S>(e, c)]; };
  auto ss = [&](int e, int c, cx_t<T> v) { tile[fx_smem<false, TILE, RX::S>(e, c)] = v; };
  PostOf<T, F> post{f, t};
  fx_stage<T, TILE, NT, RX::S, RX::S, RX::R0, false, false>(tw0, sl, ss, NoPost{});
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S0, RX::R1, false, false>(tw1, sl, ss, NoPost{});
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S1, RX::R2, false, false>(nullptr, sl, gs, post);
}

// The column DIT pass.  FROM_SMEM: the tile has already been filled (TMA, see k_fx_inv_tma) and the first stage runs
// in place on it instead of loading from HBM.
template <typename T, int TILE, int NT, class RX, class F, bool FROM_SMEM, class TRIP = NoTripHook>
__device__ __forceinline__ void fx_inv_core(unsigned char* smem_raw, const cx_t<T>* __restrict__ tw0,
                                            const cx_t<T>* __restrict__ tw1, const F& f, const int t, const int b,
                                            TRIP trip = TRIP{}) {
  cx_t<T>* tile = reinterpret_cast<cx_t<T>*>(smem_raw);
  // running max |re|, |im| of what this thread stores: the per-row max|x| the peak search needs comes for free
  T mx_re = (T)0, mx_im = (T)0;
  int ix_re = 0, ix_im = 0;  // output index of the thread's running maxima (argmax candidates)
  // Output number p of the last DIT stage owns the contiguous time block [p, p+1) * (S/R0) * N2 of the row: the
  // maxima per p are the maxima of R0 consecutive blocks of the output ("block maxima", impb_view.blk_max), with
  // which the first-peak search starts at the first block that can hold a peak instead of at sample 0.
  // (kept as the integer bit patterns of the non-negative magnitudes: integer max is cheaper than FP64 max)
  using bits_t = typename std::conditional<sizeof(T) == 8, long long, int>::type;
  auto to_bits = [](T v) -> bits_t {
    if constexpr (sizeof(T) == 8) return __double_as_longlong(v); else return __float_as_int(v);
  };
  auto from_bits = [](bits_t v) -> T {
    if constexpr (sizeof(T) == 8) return __longlong_as_double(v); else return __int_as_float(v);
  };
  bits_t bm_re[RX::R0], bm_im[RX::R0];
#pragma unroll
  for (int p = 0; p < RX::R0; ++p) bm_re[p] = bm_im[p] = 0;
  const bool blocks = f.tracks_blocks();
  auto gl = [&](int e, int c) -> cx_t<T> { return f.load(b, t, e, c); };
  auto gs = [&](int e, int c, cx_t<T> v, int p) {
    T ax, ay;
    int j;
    if (!f.store_abs(b, t, e, c, v, ax, ay, j)) return;
    if (ax > mx_re) { mx_re = ax; ix_re = j; }
    if (ay > mx_im) { mx_im = ay; ix_im = j; }
    if (blocks) { bm_re[p] = max(bm_re[p], to_bits(ax)); bm_im[p] = max(bm_im[p], to_bits(ay)); }
  };
  auto sl = [&](int e, int c) -> cx_t<T> { return tile[fx_smem<false, TILE, RX::S>(e, c)]; };
  auto ss = [&](int e, int c, cx_t<T> v) { tile[fx_smem<false, TILE, RX::S>(e, c)] = v; };
  if constexpr (FROM_SMEM) fx_stage<T, TILE, NT, RX::S, RX::S1, RX::R2, true, false>(nullptr, sl, ss, NoPost{}, trip);
  else fx_stage<T, TILE, NT, RX::S, RX::S1, RX::R2, true, false, kPrefetch>(nullptr, gl, ss, NoPost{});
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S0, RX::R1, true, false>(tw1, sl, ss, NoPost{});
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S, RX::R0, true, false>(tw0, sl, gs, NoPost{});
  if (f.tracks_max()) {
    // Everything below is per warp (integer REDUX on the bit patterns of the non-negative magnitudes): no shared
    // memory and no barrier at the end of the CTA.  Each warp publishes its maxima (atomicMax only if it beats the
    // value already there) and, for argmax, one candidate per (row, tile, warp).
    auto warp_max = [](bits_t v) -> bits_t {
      if constexpr (sizeof(T) == 8) {
        const unsigned hi = (unsigned)((unsigned long long)v >> 32);
        const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
        const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? (unsigned)(unsigned long long)v : 0u);
        return (long long)(((unsigned long long)mh << 32) | ml);
      } else {
        return (int)__reduce_max_sync(0xffffffffu, (unsigned)v);
      }
    };
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bits_t my_re = to_bits(mx_re), my_im = to_bits(mx_im);
    const bits_t w_re = warp_max(my_re), w_im = warp_max(my_im);
    if (lane == 0) f.publish_max(b, from_bits(w_re), from_bits(w_im));
    if (f.tracks_arg()) {
      const unsigned a_re = __reduce_min_sync(0xffffffffu, my_re == w_re ? (unsigned)ix_re : 0x7fffffffu);
      const unsigned a_im = __reduce_min_sync(0xffffffffu, my_im == w_im ? (unsigned)ix_im : 0x7fffffffu);
      if (lane == 0) f.publish_arg(b, t * (NT / 32) + warp, from_bits(w_re), (int)a_re, from_bits(w_im), (int)a_im);
    }
    if (blocks) {
      // warp-level integer reductions (REDUX) of the 2 * R0 bit patterns; lane q of every warp then publishes value
      // q if it beats what is already there: no shared memory, no barrier, and after the first few CTAs of a row
      // hardly any atomics
      bits_t mine = 0;
#pragma unroll
      for (int q = 0; q < 2 * RX::R0; ++q) {
        const bits_t r = warp_max((q & 1) ? bm_im[q >> 1] : bm_re[q >> 1]);
        if (lane == q) mine = r;
      }
      if (lane < 2 * RX::R0 && mine > 0) f.publish_block(b, lane >> 1, lane & 1, from_bits(mine));
    }
  }
}

template <typename T, int TILE, int NT, class RX, class F>
__global__ void __launch_bounds__(NT, 512 / NT)
k_fx_inv(const cx_t<T>* __restrict__ tw0, const cx_t<T>* __restrict__ tw1, const F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  fx_inv_core<T, TILE, NT, RX, F, false>(smem_raw, tw0, tw1, f, blockIdx.x, blockIdx.y);
}

// ---- TMA-fed variant (FP64, full-width tiles) ---------------------------------------------------------------
// The workspace tile of a column pass is S rows x 128 bytes (8 complex doubles).  One elected thread issues S/256
// cp.async.bulk.tensor.2d copies (a tensor-map box holds at most 256 rows) that land in shared memory in exactly the
// layout the butterflies use: the hardware 128-byte swizzle XORs the 16-byte chunk index with (row % 8), which IS
// slot = c ^ (e & 7).  No thread spends registers, LSU wavefronts or scoreboard waits on the HBM read.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
  const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
  unsigned ok = 0;
  int spins = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(phase)
        : "memory");
    if (!ok && ++spins > (1 << 22)) __trap();  // a copy that never lands must abort the launch, not hang the device
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
          (unsigned)__cvta_generic_to_shared(smem_dst)),
      "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"((unsigned)__cvta_generic_to_shared(bar))
      : "memory");
}

template <int TILE, int NT, class RX, class F>
__global__ void __launch_bounds__(NT, 512 / NT)
k_fx_inv_tma(const double2* __restrict__ tw0, const double2* __restrict__ tw1, const F f,
             const __grid_constant__ CUtensorMap ws_map) {
  extern __shared__ __align__(1024) unsigned char smem_tma[];
  constexpr int BOXES = RX::S / 256;
  __shared__ __align__(8) unsigned long long bar[BOXES];
  static_assert(TILE == 8 && RX::S % 256 == 0, "128-byte rows, whole 256-row boxes");
  // first DIT stage: radix R2 on consecutive rows, NT / TILE butterflies per trip = rows [it, it + 1) * R2 * NT / TILE
  static_assert(RX::R2 * NT / TILE == 256, "one trip of the first stage reads exactly one 256-row box");
  const int t = blockIdx.x, b = blockIdx.y;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < BOXES; ++k) mbar_init(&bar[k], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < BOXES; ++k) {  // coordinates in doubles (inner) and matrix rows (outer)
      mbar_expect_tx(&bar[k], 256 * TILE * (unsigned)sizeof(double2));
      tma_load_2d(smem_tma + (size_t)k * 256 * TILE * sizeof(double2), &ws_map, t * TILE * 2, b * RX::S + k * 256, &bar[k]);
    }
  }
  // trip `it` of the first stage touches only box `it`: its butterflies run while the later boxes are still in flight
  auto wait_box = [&](int it) { if (it < BOXES) mbar_wait(&bar[it], 0); };
  fx_inv_core<double, TILE, NT, RX, F, true>(smem_tma, tw0, tw1, f, t, b, wait_box);
}

// row pass: forward, x spectrum, inverse.  tw0r / tw1r are the [p-1][np] (transposed) stage tables.
// FROM_SMEM: the tile has been staged in shared memory and the first stage runs in place.  (A TMA-fed variant of this
// pass -- one 3-D box per tile, and one copy + mbarrier per matrix row -- was measured 7-8 % SLOWER than loading the
// first stage straight into registers: the pass is bound by shared-memory wavefronts and the staged tile adds a read.)
template <typename T, int TILE, int NT, class RX, class F, bool FROM_SMEM>
__device__ __forceinline__ void fx_conv_core(unsigned char* smem_raw, const cx_t<T>* __restrict__ tw0r,
                                             const cx_t<T>* __restrict__ tw1r, const F& f, const int t, const int b) {
  cx_t<T>* tile = reinterpret_cast<cx_t<T>*>(smem_raw);
  constexpr int ITEMS_MID = (RX::S / RX::R2) * TILE;
  auto gl = [&](int e, int c) -> cx_t<T> { return f.load(b, t, e, c); };
  auto gs = [&](int e, int c, cx_t<T> v) { f.store(b, t, e, c, v); };
  auto sl = [&](int e, int c) -> cx_t<T> { return tile[fx_smem<true, TILE, RX::S>(e, c)]; };
  auto ss = [&](int e, int c, cx_t<T> v) { tile[fx_smem<true, TILE, RX::S>(e, c)] = v; };
  auto pw = [&](int q, int w) -> cx_t<T> { return f.pw(t, RX::R2, ITEMS_MID, q, w); };
  PostOf<T, F> post{f, t};
  if constexpr (FROM_SMEM) fx_stage<T, TILE, NT, RX::S, RX::S, RX::R0, false, true>(tw0r, sl, ss, NoPost{});
  else fx_stage<T, TILE, NT, RX::S, RX::S, RX::R0, false, true, kPrefetch>(tw0r, gl, ss, NoPost{});
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S0, RX::R1, false, true>(tw1r, sl, ss, NoPost{});
  __syncthreads();
  fx_mid<T, TILE, NT, RX::S, RX::R2>(sl, ss, pw);
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S0, RX::R1, true, true>(tw1r, sl, ss, NoPost{});
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S, RX::R0, true, true>(tw0r, sl, gs, post);
}

template <typename T, int TILE, int NT, class RX, class F>
__global__ void __launch_bounds__(NT, 512 / NT)
k_fx_conv(const cx_t<T>* __restrict__ tw0r, const cx_t<T>* __restrict__ tw1r, const F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  fx_conv_core<T, TILE, NT, RX, F, false>(smem_raw, tw0r, tw1r, f, blockIdx.x, blockIdx.y);
}

// ------------------------------------------------------------------------------------------------
// Forward-only / inverse-only passes for plain functors (F.load / F.store, the four-step twiddle applied by the
// functor itself): the frequency-response transforms of fr.cu (K8 magnitudes, firwin2, cepstrum) run through
// these whenever their sub-FFT length has a specialised radix sequence, with full-table (exact) stage twiddles.
// ------------------------------------------------------------------------------------------------
template <typename T, int TILE, int NT, class RX, bool ROW, class F>
__global__ void __launch_bounds__(NT, 512 / NT)
k_fxg_fwd(const cx_t<T>* __restrict__ tw0, const cx_t<T>* __restrict__ tw1, const F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx_t<T>* tile = reinterpret_cast<cx_t<T>*>(smem_raw);
  const int t = blockIdx.x, b = blockIdx.y;
  auto gl = [&](int e, int c) -> cx_t<T> { return f.load(b, t, e, c); };
  auto gs = [&](int e, int c, cx_t<T> v) { f.store(b, t, e, c, v); };
  auto sl = [&](int e, int c) -> cx_t<T> { return tile[fx_smem<ROW, TILE, RX::S>(e, c)]; };
  auto ss = [&](int e, int c, cx_t<T> v) { tile[fx_smem<ROW, TILE, RX::S>(e, c)] = v; };
  fx_stage<T, TILE, NT, RX::S, RX::S, RX::R0, false, ROW, false, true>(tw0, gl, ss, NoPost{});
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S0, RX::R1, false, ROW, false, true>(tw1, sl, ss, NoPost{});
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S1, RX::R2, false, ROW, false, true>(nullptr, sl, gs, NoPost{});
}

template <typename T, int TILE, int NT, class RX, bool ROW, class F>
__global__ void __launch_bounds__(NT, 512 / NT)
k_fxg_inv(const cx_t<T>* __restrict__ tw0, const cx_t<T>* __restrict__ tw1, const F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx_t<T>* tile = reinterpret_cast<cx_t<T>*>(smem_raw);
  const int t = blockIdx.x, b = blockIdx.y;
  auto gl = [&](int e, int c) -> cx_t<T> { return f.load(b, t, e, c); };
  auto gs = [&](int e, int c, cx_t<T> v) { f.store(b, t, e, c, v); };
  auto sl = [&](int e, int c) -> cx_t<T> { return tile[fx_smem<ROW, TILE, RX::S>(e, c)]; };
  auto ss = [&](int e, int c, cx_t<T> v) { tile[fx_smem<ROW, TILE, RX::S>(e, c)] = v; };
  fx_stage<T, TILE, NT, RX::S, RX::S1, RX::R2, true, ROW, false, true>(nullptr, gl, ss, NoPost{});
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S0, RX::R1, true, ROW, false, true>(tw1, sl, ss, NoPost{});
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S, RX::R0, true, ROW, false, true>(tw0, sl, gs, NoPost{});
}

}  // namespace impb
