// Generic tile-FFT kernels.  A functor F decides what "element e of column c of tile t of batch b" means
// in HBM (column pass: stride N2; row pass: contiguous), so the same three kernels serve every
// transform in the library:
//   k_tile_fwd  : forward DIF sub-FFT        F.load -> F.store (position order)
//   k_tile_inv  : inverse DIT sub-FFT        F.load (position order) -> F.store
//   k_tile_conv : forward, x pointwise, inverse on the same tile (F.load -> F.pw -> F.store)
// grid = (tiles, batch); block = plan.threads; dynamic smem = tile_smem_bytes().
#pragma once
#include "fft_tile.cuh"

namespace impb {

template <typename T, int TILE, bool ROW, class F>
__global__ void __launch_bounds__(512) k_tile_fwd(const SubPlan sp, const cx_t<T>* __restrict__ tab_g, const F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx_t<T>* tile = reinterpret_cast<cx_t<T>*>(smem_raw);
  cx_t<T>* tab = tile + (size_t)tile_row_pitch(sp.S) * TILE;
  load_table<T>(tab, tab_g, sp.ntab);
  __syncthreads();
  const int t = blockIdx.x, b = blockIdx.y;
  auto gl = [&](int e, int c) -> cx_t<T> { return f.load(b, t, e, c); };
  auto gs = [&](int e, int c, cx_t<T> v) { f.store(b, t, e, c, v); };
  tile_fft_forward<T, TILE, ROW>(sp, tile, tab, gl, gs);
}

template <typename T, int TILE, bool ROW, class F>
__global__ void __launch_bounds__(512) k_tile_inv(const SubPlan sp, const cx_t<T>* __restrict__ tab_g, const F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx_t<T>* tile = reinterpret_cast<cx_t<T>*>(smem_raw);
  cx_t<T>* tab = tile + (size_t)tile_row_pitch(sp.S) * TILE;
  load_table<T>(tab, tab_g, sp.ntab);
  __syncthreads();
  const int t = blockIdx.x, b = blockIdx.y;
  auto gl = [&](int e, int c) -> cx_t<T> { return f.load(b, t, e, c); };
  auto gs = [&](int e, int c, cx_t<T> v) { f.store(b, t, e, c, v); };
  tile_fft_inverse<T, TILE, ROW>(sp, tile, tab, gl, gs);
}

template <typename T, int TILE, class F>
__global__ void __launch_bounds__(512) k_tile_conv(const SubPlan sp, const cx_t<T>* __restrict__ tab_g, const F f) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx_t<T>* tile = reinterpret_cast<cx_t<T>*>(smem_raw);
  cx_t<T>* tab = tile + (size_t)tile_row_pitch(sp.S) * TILE;
  load_table<T>(tab, tab_g, sp.ntab);
  __syncthreads();
  const int t = blockIdx.x, b = blockIdx.y;
  const int R = sp.radix[sp.nst - 1];
  const int items = (sp.S / R) * TILE;
  auto gl = [&](int e, int c) -> cx_t<T> { return f.load(b, t, e, c); };
  auto gs = [&](int e, int c, cx_t<T> v) { f.store(b, t, e, c, v); };
  auto pw = [&](int q, int w) -> cx_t<T> { return f.pw(t, R, items, q, w); };
  tile_fft_conv<T, TILE>(sp, tile, tab, gl, gs, pw);
}

// ---- plain complex [batch][N] functors (N = N1*N2 row-major [N1][N2]) -------------------------------
// column pass: tile t = columns [t*TILE, ..), element e = row index
template <typename T, int TILE> struct ColPlain {
  const cx_t<T>* src;
  cx_t<T>* dst;
  const cx_t<T>* tw;  // nullptr: none; else multiplied on store (conj if conj_tw)
  int64_t N;
  int N2;
  int conj_tw;
  __device__ __forceinline__ cx_t<T> load(int b, int t, int e, int c) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return mkc<T>(0, 0);
    return src[(size_t)b * N + (size_t)e * N2 + n2];
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, cx_t<T> v) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return;
    const size_t idx = (size_t)e * N2 + n2;
    if (tw) v = cmul(v, conj_tw ? cconj(tw[idx]) : tw[idx]);
    dst[(size_t)b * N + idx] = v;
  }
};

// row pass: tile t = rows [t*TILE, ..), element e = column index
template <typename T, int TILE> struct RowPlain {
  const cx_t<T>* src;
  cx_t<T>* dst;
  const cx_t<T>* tw;
  int64_t N;
  int N1, N2;
  int conj_tw;
  int tw_on_load;
  __device__ __forceinline__ cx_t<T> load(int b, int t, int e, int c) const {
    const int row = t * TILE + c;
    if (row >= N1) return mkc<T>(0, 0);
    const size_t idx = (size_t)row * N2 + e;
    cx_t<T> v = src[(size_t)b * N + idx];
    if (tw && tw_on_load) v = cmul(v, conj_tw ? cconj(tw[idx]) : tw[idx]);
    return v;
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, cx_t<T> v) const {
    const int row = t * TILE + c;
    if (row >= N1) return;
    const size_t idx = (size_t)row * N2 + e;
    if (tw && !tw_on_load) v = cmul(v, conj_tw ? cconj(tw[idx]) : tw[idx]);
    dst[(size_t)b * N + idx] = v;
  }
};

template <typename K> inline cudaError_t set_smem(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace impb
