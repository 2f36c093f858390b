// Host-side handle for one N = N1 x N2 transform: owns the device tables of a Plan2D and launches passes.
#pragma once
#include <vector>

#include "common.cuh"
#include "fft_kernels.cuh"

namespace impb {

template <typename T> struct Fft2D {
  Plan2D p;
  cx_t<T>* tw = nullptr;       // [N1][N2]: exp(-2 pi i n2 k1(p1) / N), k1 = perm_col[p1]
  cx_t<T>* tab_col = nullptr;  // sub-FFT twiddle tables
  cx_t<T>* tab_row = nullptr;
  cx_t<T>* fx_col[2] = {nullptr, nullptr};  // per-stage tables of the specialised kernels (if fixed)
  cx_t<T>* fx_row[2] = {nullptr, nullptr};
  bool col_fixed = false, row_fixed = false;
  int* perm_col = nullptr;     // device copies of the digit-reversal maps
  int* perm_row = nullptr;
  std::vector<int> h_perm_col, h_perm_row;

  int init(const Plan2D& plan) {
    p = plan;
    h_perm_col = subplan_perm(p.col);
    h_perm_row = subplan_perm(p.row);
    const size_t N = (size_t)p.N;
    std::vector<cx_t<T>> h(N);
    for (int p1 = 0; p1 < p.N1; ++p1) {
      const int64_t k1 = h_perm_col[p1];
      for (int n2 = 0; n2 < p.N2; ++n2) {
        double re, im;
        unit_root((k1 * n2) % p.N, p.N, &re, &im);
        h[(size_t)p1 * p.N2 + n2] = mkc<T>((T)re, (T)im);
      }
    }
    IMPB_CUDA(cudaMalloc(&tw, N * sizeof(cx_t<T>)));
    IMPB_CUDA(cudaMemcpy(tw, h.data(), N * sizeof(cx_t<T>), cudaMemcpyHostToDevice));
    for (int which = 0; which < 2; ++which) {
      const SubPlan& sp = which ? p.row : p.col;
      std::vector<double> t = subplan_table(sp);
      std::vector<cx_t<T>> tt(sp.ntab);
      for (int i = 0; i < sp.ntab; ++i) tt[i] = mkc<T>((T)t[2 * i], (T)t[2 * i + 1]);
      cx_t<T>** dst = which ? &tab_row : &tab_col;
      IMPB_CUDA(cudaMalloc(dst, sp.ntab * sizeof(cx_t<T>)));
      IMPB_CUDA(cudaMemcpy(*dst, tt.data(), sp.ntab * sizeof(cx_t<T>), cudaMemcpyHostToDevice));
    }
    for (int which = 0; which < 2; ++which) {
      const SubPlan& sp = which ? p.row : p.col;
      int fr[3];
      const bool fixed = fixed_radices(sp.S, fr) && sp.nst == 3;
      (which ? row_fixed : col_fixed) = fixed;
      if (!fixed) continue;
      for (int s = 0; s < 2; ++s) {
        std::vector<double> t = stage_table(sp, s, /*transposed=*/which == 1);
        std::vector<cx_t<T>> tt(t.size() / 2);
        for (size_t i = 0; i < tt.size(); ++i) tt[i] = mkc<T>((T)t[2 * i], (T)t[2 * i + 1]);
        cx_t<T>** dst = which ? &fx_row[s] : &fx_col[s];
        IMPB_CUDA(cudaMalloc(dst, tt.size() * sizeof(cx_t<T>)));
        IMPB_CUDA(cudaMemcpy(*dst, tt.data(), tt.size() * sizeof(cx_t<T>), cudaMemcpyHostToDevice));
      }
    }
    IMPB_CUDA(cudaMalloc(&perm_col, p.N1 * sizeof(int)));
    IMPB_CUDA(cudaMalloc(&perm_row, p.N2 * sizeof(int)));
    IMPB_CUDA(cudaMemcpy(perm_col, h_perm_col.data(), p.N1 * sizeof(int), cudaMemcpyHostToDevice));
    IMPB_CUDA(cudaMemcpy(perm_row, h_perm_row.data(), p.N2 * sizeof(int), cudaMemcpyHostToDevice));
    return IMPB_OK;
  }

  void release() {
    cudaFree(tw); cudaFree(tab_col); cudaFree(tab_row); cudaFree(perm_col); cudaFree(perm_row);
    for (int s = 0; s < 2; ++s) { cudaFree(fx_col[s]); cudaFree(fx_row[s]); fx_col[s] = fx_row[s] = nullptr; }
    tw = tab_col = tab_row = nullptr;
    perm_col = perm_row = nullptr;
  }

  int col_tiles() const { return (p.N2 + p.tile - 1) / p.tile; }
  int row_tiles() const { return (p.N1 + p.tile - 1) / p.tile; }

  // launch helpers; F is a functor type, TILE must equal p.tile
  template <int TILE, class F> int col_fwd(const F& f, int batch, cudaStream_t st) const {
    auto k = k_tile_fwd<T, TILE, false, F>;
    IMPB_CUDA(set_smem(k, p.smem_col));
    k<<<dim3(col_tiles(), batch), p.threads, p.smem_col, st>>>(p.col, tab_col, f);
    IMPB_CUDA(cudaGetLastError());
    return IMPB_OK;
  }
  template <int TILE, class F> int col_inv(const F& f, int batch, cudaStream_t st) const {
    auto k = k_tile_inv<T, TILE, false, F>;
    IMPB_CUDA(set_smem(k, p.smem_col));
    k<<<dim3(col_tiles(), batch), p.threads, p.smem_col, st>>>(p.col, tab_col, f);
    IMPB_CUDA(cudaGetLastError());
    return IMPB_OK;
  }
  template <int TILE, class F> int row_fwd(const F& f, int batch, cudaStream_t st) const {
    auto k = k_tile_fwd<T, TILE, true, F>;
    IMPB_CUDA(set_smem(k, p.smem_row));
    k<<<dim3(row_tiles(), batch), p.threads, p.smem_row, st>>>(p.row, tab_row, f);
    IMPB_CUDA(cudaGetLastError());
    return IMPB_OK;
  }
  template <int TILE, class F> int row_inv(const F& f, int batch, cudaStream_t st) const {
    auto k = k_tile_inv<T, TILE, true, F>;
    IMPB_CUDA(set_smem(k, p.smem_row));
    k<<<dim3(row_tiles(), batch), p.threads, p.smem_row, st>>>(p.row, tab_row, f);
    IMPB_CUDA(cudaGetLastError());
    return IMPB_OK;
  }
  template <int TILE, class F> int row_conv(const F& f, int batch, cudaStream_t st) const {
    auto k = k_tile_conv<T, TILE, F>;
    IMPB_CUDA(set_smem(k, p.smem_row));
    k<<<dim3(row_tiles(), batch), p.threads, p.smem_row, st>>>(p.row, tab_row, f);
    IMPB_CUDA(cudaGetLastError());
    return IMPB_OK;
  }
};

}  // namespace impb
