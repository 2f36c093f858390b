// Host-side handle for one N = N1 x N2 transform: owns the device tables of a Plan2D and launches passes.
#pragma once
#include <mutex>
#include <set>
#include <stdlib.h>
#include <vector>

#include "common.cuh"
#include "fft_fixed.cuh"
#include "fft_kernels.cuh"

namespace impb {

template <typename T> struct Fft2D {
  Plan2D p;
  cx_t<T>* tw = nullptr;       // [N1][N2]: exp(-2 pi i n2 k1(p1) / N), k1 = perm_col[p1]
  cx_t<T>* tab_col = nullptr;  // sub-FFT twiddle tables
  cx_t<T>* tab_row = nullptr;
  cx_t<T>* fx_col[2] = {nullptr, nullptr};  // per-stage tables of the specialised kernels (if fixed)
  cx_t<T>* fx_row[2] = {nullptr, nullptr};
  cx_t<T>* fxf_col[2] = {nullptr, nullptr};  // full (exact) stage tables of the plain specialised passes
  cx_t<T>* fxf_row[2] = {nullptr, nullptr};
  int device = 0;
  bool col_fixed = false, row_fixed = false;
  int* perm_col = nullptr;     // device copies of the digit-reversal maps
  int* perm_row = nullptr;
  cx_t<T>* dcol = nullptr;     // [N2] exp(-2 pi i n2 (N1 / R_last_col) / N): four-step twiddle step of the column pass
  cx_t<T>* drow = nullptr;     // [N1] exp(-2 pi i k1(p1) (N2 / R_first_row) / N): ... of the row pass (by position p1)
  int* inv_col = nullptr;      // their inverses (natural index -> position): mirror-bin lookups of paired real rows
  int* inv_row = nullptr;
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
      const bool fixed = fixed_radices(sp.S, fr) && sp.nst == 3 && fixed_tile_ok(sp.S, p.tile == TileOf<T>::value);
      (which ? row_fixed : col_fixed) = fixed;
      if (!fixed) continue;
      for (int s = 0; s < 2; ++s) {
        std::vector<double> t = stage_table(sp, s, /*transposed=*/which == 1);
        std::vector<cx_t<T>> tt(t.size() / 2);
        for (size_t i = 0; i < tt.size(); ++i) tt[i] = mkc<T>((T)t[2 * i], (T)t[2 * i + 1]);
        cx_t<T>** dst = which ? &fx_row[s] : &fx_col[s];
        IMPB_CUDA(cudaMalloc(dst, tt.size() * sizeof(cx_t<T>)));
        IMPB_CUDA(cudaMemcpy(*dst, tt.data(), tt.size() * sizeof(cx_t<T>), cudaMemcpyHostToDevice));
        std::vector<double> tf = stage_table_full(sp, s, /*transposed=*/which == 1);
        std::vector<cx_t<T>> ttf(tf.size() / 2);
        for (size_t i = 0; i < ttf.size(); ++i) ttf[i] = mkc<T>((T)tf[2 * i], (T)tf[2 * i + 1]);
        dst = which ? &fxf_row[s] : &fxf_col[s];
        IMPB_CUDA(cudaMalloc(dst, ttf.size() * sizeof(cx_t<T>)));
        IMPB_CUDA(cudaMemcpy(*dst, ttf.data(), ttf.size() * sizeof(cx_t<T>), cudaMemcpyHostToDevice));
      }
    }
    if (col_fixed && row_fixed) {
      // progression form of the four-step twiddle (fft_fixed.cuh: PostOf): output p of the last column-DIF stage sits
      // at frequency k_base + p * N1 / R_last, output p of the last row-DIT stage at sample np + p * N2 / R_first
      int fc[3], frw[3];
      fixed_radices(p.N1, fc);
      fixed_radices(p.N2, frw);
      const int64_t inc_col = p.N1 / fc[2], inc_row = p.N2 / frw[0];
      std::vector<cx_t<T>> hc((size_t)p.N2), hr((size_t)p.N1);
      for (int n2 = 0; n2 < p.N2; ++n2) {
        double re, im;
        unit_root(((int64_t)n2 * inc_col) % p.N, p.N, &re, &im);
        hc[(size_t)n2] = mkc<T>((T)re, (T)im);
      }
      for (int p1 = 0; p1 < p.N1; ++p1) {
        double re, im;
        unit_root(((int64_t)h_perm_col[p1] * inc_row) % p.N, p.N, &re, &im);
        hr[(size_t)p1] = mkc<T>((T)re, (T)im);
      }
      IMPB_CUDA(cudaMalloc(&dcol, p.N2 * sizeof(cx_t<T>)));
      IMPB_CUDA(cudaMalloc(&drow, p.N1 * sizeof(cx_t<T>)));
      IMPB_CUDA(cudaMemcpy(dcol, hc.data(), p.N2 * sizeof(cx_t<T>), cudaMemcpyHostToDevice));
      IMPB_CUDA(cudaMemcpy(drow, hr.data(), p.N1 * sizeof(cx_t<T>), cudaMemcpyHostToDevice));
    }
    IMPB_CUDA(cudaGetDevice(&device));
    IMPB_CUDA(cudaMalloc(&perm_col, p.N1 * sizeof(int)));
    IMPB_CUDA(cudaMalloc(&perm_row, p.N2 * sizeof(int)));
    IMPB_CUDA(cudaMemcpy(perm_col, h_perm_col.data(), p.N1 * sizeof(int), cudaMemcpyHostToDevice));
    IMPB_CUDA(cudaMemcpy(perm_row, h_perm_row.data(), p.N2 * sizeof(int), cudaMemcpyHostToDevice));
    std::vector<int> ic((size_t)p.N1), ir((size_t)p.N2);
    for (int i = 0; i < p.N1; ++i) ic[(size_t)h_perm_col[i]] = i;
    for (int i = 0; i < p.N2; ++i) ir[(size_t)h_perm_row[i]] = i;
    IMPB_CUDA(cudaMalloc(&inv_col, p.N1 * sizeof(int)));
    IMPB_CUDA(cudaMalloc(&inv_row, p.N2 * sizeof(int)));
    IMPB_CUDA(cudaMemcpy(inv_col, ic.data(), p.N1 * sizeof(int), cudaMemcpyHostToDevice));
    IMPB_CUDA(cudaMemcpy(inv_row, ir.data(), p.N2 * sizeof(int), cudaMemcpyHostToDevice));
    return IMPB_OK;
  }

  void release() {
    cudaFree(tw); cudaFree(tab_col); cudaFree(tab_row); cudaFree(perm_col); cudaFree(perm_row);
    cudaFree(inv_col); cudaFree(inv_row); cudaFree(dcol); cudaFree(drow);
    inv_col = inv_row = nullptr;
    dcol = drow = nullptr;
    for (int s = 0; s < 2; ++s) {
      cudaFree(fx_col[s]); cudaFree(fx_row[s]); cudaFree(fxf_col[s]); cudaFree(fxf_row[s]);
      fx_col[s] = fx_row[s] = fxf_col[s] = fxf_row[s] = nullptr;
    }
    tw = tab_col = tab_row = nullptr;
    perm_col = perm_row = nullptr;
  }

  int col_tiles() const { return (p.N2 + p.tile - 1) / p.tile; }
  int row_tiles() const { return (p.N1 + p.tile - 1) / p.tile; }

  // ---- specialised plain passes (fft_fixed.cuh:k_fxg_*): lengths of plan.cu:fixed_radices, full-width tiles ----------
  template <class K> static int fxg_prepare(K kernel, size_t smem, int device) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    const std::pair<const void*, int> key((const void*)kernel, device);
    std::lock_guard<std::mutex> lk(mu);
    if (done.count(key)) return IMPB_OK;
    IMPB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done.insert(key);
    return IMPB_OK;
  }
// short sub-FFTs (S <= 192: 15-24 butterflies per column and stage) leave half of a 256-thread CTA idle: 128 threads
#define IMPB_FXG_CASE(S_, R0_, R1_, R2_)                                                             \
  case S_: {                                                                                         \
    constexpr int NTX = (S_ <= 192) ? 128 : 256;                                                      \
    auto k = INV ? k_fxg_inv<T, TILE, NTX, Radix3<R0_, R1_, R2_>, ROW, F>                             \
                 : k_fxg_fwd<T, TILE, NTX, Radix3<R0_, R1_, R2_>, ROW, F>;                            \
    const size_t smem = fx_tile_elems<S_, TILE>() * sizeof(cx_t<T>);                                  \
    int rc = fxg_prepare(k, smem, device);                                                            \
    if (rc) return rc;                                                                                \
    k<<<grid, NTX, smem, st>>>(tw0, tw1, f);                                                          \
    IMPB_CUDA(cudaGetLastError());                                                                    \
    return IMPB_OK;                                                                                   \
  }
  template <int TILE, bool ROW, bool INV, class F> int fxg_launch(int S, dim3 grid, const F& f, cudaStream_t st) const {
    const cx_t<T>* tw0 = ROW ? fxf_row[0] : fxf_col[0];
    const cx_t<T>* tw1 = ROW ? fxf_row[1] : fxf_col[1];
    switch (S) {
      IMPB_FXG_CASE(768, 8, 12, 8)
      IMPB_FXG_CASE(675, 15, 9, 5)
      IMPB_FXG_CASE(432, 12, 9, 4)
      IMPB_FXG_CASE(400, 10, 8, 5)
      IMPB_FXG_CASE(300, 12, 5, 5)
      IMPB_FXG_CASE(256, 8, 8, 4)
      IMPB_FXG_CASE(243, 9, 9, 3)
      IMPB_FXG_CASE(162, 9, 6, 3)
      IMPB_FXG_CASE(160, 8, 5, 4)
      IMPB_FXG_CASE(135, 5, 9, 3)
      IMPB_FXG_CASE(128, 8, 4, 4)
      IMPB_FXG_CASE(125, 5, 5, 5)
      IMPB_FXG_CASE(120, 8, 5, 3)
      IMPB_FXG_CASE(108, 4, 9, 3)
      IMPB_FXG_CASE(81, 3, 9, 3)
      IMPB_FXG_CASE(75, 5, 5, 3)
      default: set_error("no specialised pass for S=%d", S); return IMPB_EUNSUPPORTED;
    }
  }
#undef IMPB_FXG_CASE
  template <int TILE> bool fxg_ok(bool row) const {
    static const bool off = getenv("IMPB_NO_FIXED") != nullptr;
    return !off && TILE == TileOf<T>::value && (row ? row_fixed : col_fixed);
  }

  // launch helpers; F is a functor type, TILE must equal p.tile
  template <int TILE, bool ALLOW_FX = true, class F> int col_fwd(const F& f, int batch, cudaStream_t st) const {
    if constexpr (ALLOW_FX) if (fxg_ok<TILE>(false)) return fxg_launch<TILE, false, false>(p.N1, dim3(col_tiles(), batch), f, st);
    auto k = k_tile_fwd<T, TILE, false, F>;
    IMPB_CUDA(set_smem(k, p.smem_col));
    k<<<dim3(col_tiles(), batch), p.threads, p.smem_col, st>>>(p.col, tab_col, f);
    IMPB_CUDA(cudaGetLastError());
    return IMPB_OK;
  }
  template <int TILE, bool ALLOW_FX = true, class F> int col_inv(const F& f, int batch, cudaStream_t st) const {
    if constexpr (ALLOW_FX) if (fxg_ok<TILE>(false)) return fxg_launch<TILE, false, true>(p.N1, dim3(col_tiles(), batch), f, st);
    auto k = k_tile_inv<T, TILE, false, F>;
    IMPB_CUDA(set_smem(k, p.smem_col));
    k<<<dim3(col_tiles(), batch), p.threads, p.smem_col, st>>>(p.col, tab_col, f);
    IMPB_CUDA(cudaGetLastError());
    return IMPB_OK;
  }
  template <int TILE, bool ALLOW_FX = true, class F> int row_fwd(const F& f, int batch, cudaStream_t st) const {
    if constexpr (ALLOW_FX) if (fxg_ok<TILE>(true)) return fxg_launch<TILE, true, false>(p.N2, dim3(row_tiles(), batch), f, st);
    auto k = k_tile_fwd<T, TILE, true, F>;
    IMPB_CUDA(set_smem(k, p.smem_row));
    k<<<dim3(row_tiles(), batch), p.threads, p.smem_row, st>>>(p.row, tab_row, f);
    IMPB_CUDA(cudaGetLastError());
    return IMPB_OK;
  }
  template <int TILE, bool ALLOW_FX = true, class F> int row_inv(const F& f, int batch, cudaStream_t st) const {
    if constexpr (ALLOW_FX) if (fxg_ok<TILE>(true)) return fxg_launch<TILE, true, true>(p.N2, dim3(row_tiles(), batch), f, st);
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
