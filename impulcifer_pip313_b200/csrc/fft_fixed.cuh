// Compile-time specialised tile FFTs: same algorithm and memory layout as fft_tile.cuh, but the radix
// sequence <R0,R1,R2> is a template parameter, so every stride, shared-memory offset and loop trip count
// is an immediate.  Inter-stage twiddles come from small per-stage tables tw_s[np][p-1] (read through the
// read-only path; the 8 columns of a tile share each entry, so a warp touches one or two sectors).
// Used for the sub-FFT lengths the headline workloads need; every other length takes the generic path.
#pragma once
#include "fft_tile.cuh"

namespace impb {

template <int R0_, int R1_, int R2_> struct Radix3 {
  static constexpr int R0 = R0_, R1 = R1_, R2 = R2_;
  static constexpr int S = R0_ * R1_ * R2_;
  static constexpr int S0 = S / R0_;   // stride of stage 0 == length entering stage 1
  static constexpr int S1 = S0 / R1_;  // stride of stage 1 == R2
  static constexpr int TW0 = S0 * (R0_ - 1);  // entries of the stage-0 table
  static constexpr int TW1 = S1 * (R1_ - 1);
};

// one radix-R stage over the whole tile; LD/ST address elements as (e, c)
template <typename T, int TILE, int NT, int S, int LEN, int R, bool INV, class LD, class ST>
__device__ __forceinline__ void fx_stage(const cx_t<T>* __restrict__ tw, LD ld, ST st) {
  constexpr int STRIDE = LEN / R;
  constexpr int ITEMS = (S / R) * TILE;
  constexpr int TRIPS = (ITEMS + NT - 1) / NT;
#pragma unroll 1
  for (int it = 0; it < TRIPS; ++it) {
    const int w = threadIdx.x + it * NT;
    if ((ITEMS % NT) != 0 && w >= ITEMS) break;
    const int c = w % TILE, j = w / TILE;
    const int b = j / STRIDE, np = j % STRIDE;
    const int base = b * LEN + np;
    cx_t<T> x[R];
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = ld(base + q * STRIDE, c);
    if (STRIDE > 1) {
      const cx_t<T>* twp = tw + np * (R - 1);
      if (INV) {
#pragma unroll
        for (int p = 1; p < R; ++p) x[p] = cmul(x[p], cconj(__ldg(twp + p - 1)));
        Dft<T, R, true>::run(x);
      } else {
        Dft<T, R, false>::run(x);
#pragma unroll
        for (int p = 1; p < R; ++p) x[p] = cmul(x[p], __ldg(twp + p - 1));
      }
    } else {
      Dft<T, R, INV>::run(x);
    }
#pragma unroll
    for (int p = 0; p < R; ++p) st(base + p * STRIDE, c, x[p]);
  }
}

// last forward stage x pointwise x first inverse stage, in registers
template <typename T, int TILE, int NT, int S, int R, class LD, class ST, class PW>
__device__ __forceinline__ void fx_mid(LD ld, ST st, PW pw) {
  constexpr int ITEMS = (S / R) * TILE;
  constexpr int TRIPS = (ITEMS + NT - 1) / NT;
#pragma unroll 1
  for (int it = 0; it < TRIPS; ++it) {
    const int w = threadIdx.x + it * NT;
    if ((ITEMS % NT) != 0 && w >= ITEMS) break;
    const int c = w % TILE, j = w / TILE;
    cx_t<T> x[R];
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = ld(j * R + q, c);
    Dft<T, R, false>::run(x);
#pragma unroll
    for (int q = 0; q < R; ++q) x[q] = cmul(x[q], pw(q, w));
    Dft<T, R, true>::run(x);
#pragma unroll
    for (int q = 0; q < R; ++q) st(j * R + q, c, x[q]);
  }
}

template <typename T, int TILE, int NT, class RX, class F>
__global__ void __launch_bounds__(NT, 2)
k_fx_fwd(const cx_t<T>* __restrict__ tw0, const cx_t<T>* __restrict__ tw1, const F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx_t<T>* tile = reinterpret_cast<cx_t<T>*>(smem_raw);
  constexpr int PITCH = TILE + 1;
  const int t = blockIdx.x, b = blockIdx.y;
  auto gl = [&](int e, int c) -> cx_t<T> { return f.load(b, t, e, c); };
  auto gs = [&](int e, int c, cx_t<T> v) { f.store(b, t, e, c, v); };
  auto sl = [&](int e, int c) -> cx_t<T> { return tile[e * PITCH + c]; };
  auto ss = [&](int e, int c, cx_t<T> v) { tile[e * PITCH + c] = v; };
  fx_stage<T, TILE, NT, RX::S, RX::S, RX::R0, false>(tw0, gl, ss);
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S0, RX::R1, false>(tw1, sl, ss);
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S1, RX::R2, false>(nullptr, sl, gs);
}

template <typename T, int TILE, int NT, class RX, class F>
__global__ void __launch_bounds__(NT, 2)
k_fx_inv(const cx_t<T>* __restrict__ tw0, const cx_t<T>* __restrict__ tw1, const F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx_t<T>* tile = reinterpret_cast<cx_t<T>*>(smem_raw);
  constexpr int PITCH = TILE + 1;
  const int t = blockIdx.x, b = blockIdx.y;
  auto gl = [&](int e, int c) -> cx_t<T> { return f.load(b, t, e, c); };
  auto gs = [&](int e, int c, cx_t<T> v) { f.store(b, t, e, c, v); };
  auto sl = [&](int e, int c) -> cx_t<T> { return tile[e * PITCH + c]; };
  auto ss = [&](int e, int c, cx_t<T> v) { tile[e * PITCH + c] = v; };
  fx_stage<T, TILE, NT, RX::S, RX::S1, RX::R2, true>(nullptr, gl, ss);
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S0, RX::R1, true>(tw1, sl, ss);
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S, RX::R0, true>(tw0, sl, gs);
}

template <typename T, int TILE, int NT, class RX, class F>
__global__ void __launch_bounds__(NT, 2)
k_fx_conv(const cx_t<T>* __restrict__ tw0, const cx_t<T>* __restrict__ tw1, const F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx_t<T>* tile = reinterpret_cast<cx_t<T>*>(smem_raw);
  constexpr int PITCH = TILE + 1;
  const int t = blockIdx.x, b = blockIdx.y;
  constexpr int ITEMS_MID = (RX::S / RX::R2) * TILE;
  auto gl = [&](int e, int c) -> cx_t<T> { return f.load(b, t, e, c); };
  auto gs = [&](int e, int c, cx_t<T> v) { f.store(b, t, e, c, v); };
  auto sl = [&](int e, int c) -> cx_t<T> { return tile[e * PITCH + c]; };
  auto ss = [&](int e, int c, cx_t<T> v) { tile[e * PITCH + c] = v; };
  auto pw = [&](int q, int w) -> cx_t<T> { return f.pw(t, RX::R2, ITEMS_MID, q, w); };
  fx_stage<T, TILE, NT, RX::S, RX::S, RX::R0, false>(tw0, gl, ss);
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S0, RX::R1, false>(tw1, sl, ss);
  __syncthreads();
  fx_mid<T, TILE, NT, RX::S, RX::R2>(sl, ss, pw);
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S0, RX::R1, true>(tw1, sl, ss);
  __syncthreads();
  fx_stage<T, TILE, NT, RX::S, RX::S, RX::R0, true>(tw0, sl, gs);
}

}  // namespace impb
