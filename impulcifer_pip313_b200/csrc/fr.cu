// Per-IR frequency-response post-processing on the device.
//   impb_magnitude_db   audio_io.magnitude_response                      core/audio_io.py:100-113
//   impb_magnitude_lin  abs(rfft(seg, n=nfft)) of the direct-sound windows    core/microphone_deviation_correction.py:130-133
//   impb_loglerp        InterpolatedUnivariateSpline(log10 f, y, k=1)    autoeq/frequency_response.py:874-894
//   impb_savgol_blend   FrequencyResponse._smoothen_fractional_octave    autoeq/frequency_response.py:1060-1105
//   impb_minphase_fir   firwin2 + minimum_phase(homomorphic, half=True)  autoeq/frequency_response.py:672-680,
//                       scipy/signal/_fir_filter_design.py:577-775, 1383-1406
// The transforms reuse the tile-FFT engine (fft2d.cuh) with exact-length plans; spectra stay in position
// ("scrambled") order between a forward and an inverse transform, so no reordering pass exists.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "fft2d.cuh"

using namespace impb;

struct impb_fft {
  impb_ctx* ctx = nullptr;
  int64_t n = 0;
  Fft2D<double> f;
  // lengths with a prime factor > 5 (e.g. the 391 270 samples of an uncropped 48 kHz response, which the reference
  // hands to frequency_response() in its headphone-compensation stage): Bluestein.  X[k] = conj(w_k) sum_n x[n] conj(w_n)
  // w_(k-n) with w_m = exp(i pi m^2 / n): a circular convolution of length M >= 2n - 1 (5-smooth, plan `f`), and
  // |X[k]| needs no final chirp multiply.
  bool chirp = false;
  int64_t M = 0;
  double2* w = nullptr;   // [n]  exp(+i pi m^2 / n)
  double2* Hs = nullptr;  // [M]  spectrum of the circularly extended chirp, in position order
  // folded transforms (impb_magnitude_db_folded): per (first, step) served by this plan,
  // [n] exp(-2 pi i first m / (n step)) followed by [step] exp(-2 pi i first r / step); built once, freed with the plan
  std::map<std::pair<int64_t, int64_t>, double2*> fold_w;
  std::mutex fold_mu;
};

namespace {

constexpr int TILE = 8;

// ---- real rows in (zero imaginary part), column pass ------------------------------------------------
struct RealColIn {
  const double* x;   // [batch, ld]
  int64_t ld, n;     // valid samples per row (zero-padded to N)
  double2* ws;
  const double2* tw;
  int64_t N;
  int N2;
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int n2 = t * TILE + c;
    const int64_t i = (int64_t)e * N2 + n2;
    if (n2 >= N2 || i >= n) return make_double2(0.0, 0.0);
    return make_double2(x[(size_t)b * ld + i], 0.0);
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return;
    const size_t idx = (size_t)e * N2 + n2;
    ws[(size_t)b * N + idx] = cmul(v, tw[idx]);
  }
};

// ---- row pass whose store sees the natural frequency index k ------------------------------------------
template <class OP> struct RowFwdOp {
  const double2* ws;
  int64_t N;
  int N1, N2;
  const int* perm1;
  const int* perm2;
  OP op;
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int row = t * TILE + c;
    if (row >= N1) return make_double2(0.0, 0.0);
    return ws[(size_t)b * N + (size_t)row * N2 + e];
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int row = t * TILE + c;
    if (row >= N1) return;
    op(b, (size_t)row * N2 + e, (int64_t)perm1[row] + (int64_t)N1 * perm2[e], v);
  }
};

struct MagDbOp {
  double* out;
  int64_t ld, half;
  __device__ __forceinline__ void operator()(int b, size_t, int64_t k, double2 v) const {
    if (k < half) out[(size_t)b * ld + k] = 20.0 * log10(hypot(v.x, v.y));
  }
};

// Bluestein: modulated real row in, |y[k]| in dB out
struct ChirpColIn {
  const double* x;
  int64_t ld, n;
  const double2* w;
  double2* ws;
  const double2* tw;
  int64_t N;
  int N2;
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int n2 = t * TILE + c;
    const int64_t i = (int64_t)e * N2 + n2;
    if (n2 >= N2 || i >= n) return make_double2(0.0, 0.0);
    const double v = x[(size_t)b * ld + i];
    const double2 ww = w[i];
    return make_double2(v * ww.x, -v * ww.y);  // x[n] * conj(w_n)
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return;
    const size_t idx = (size_t)e * N2 + n2;
    ws[(size_t)b * N + idx] = cmul(v, tw[idx]);
  }
};
struct ChirpColOut {
  const double2* ws;
  int64_t N;
  int N2;
  double* out;
  int64_t ld, half;
  double scale;  // 1 / M of the unnormalised inverse transform
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return make_double2(0.0, 0.0);
    return ws[(size_t)b * N + (size_t)e * N2 + n2];
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int n2 = t * TILE + c;
    const int64_t k = (int64_t)e * N2 + n2;
    if (n2 >= N2 || k >= half) return;
    out[(size_t)b * ld + k] = 20.0 * log10(hypot(v.x, v.y) * scale);
  }
};
__global__ void __launch_bounds__(256) k_cmul_rows(double2* __restrict__ ws, const double2* __restrict__ h, int64_t N) {
  double2* row = ws + (size_t)blockIdx.y * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    row[i] = cmul(row[i], h[i]);
}

struct MagLinOp {  // |X[k]| for k < nbins (natural order)
  double* out;
  int64_t ld, nbins;
  __device__ __forceinline__ void operator()(int b, size_t, int64_t k, double2 v) const {
    if (k < nbins) out[(size_t)b * ld + k] = hypot(v.x, v.y);
  }
};

// Bins first + step*j of the DFT of a real row of n_full = step * M samples are the M-point DFT of the folded, modulated row
//   y[m] = W^(first m) * sum_r x[m + M r] * exp(-2 pi i first r / step),  W = exp(-2 pi i / n_full):
// the column pass of the M-point plan folds on load (step strided, coalesced reads) and the transform is step times smaller.
struct FoldColIn {
  const double* x;
  int64_t ld;
  const int64_t* offs;  // optional row starts (elements from x)
  int64_t M;
  int step;
  const double2* wm;  // [M] W^(first m)
  const double2* wr;  // [step]
  double2* ws;
  const double2* tw;
  int64_t N;
  int N2;
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int n2 = t * TILE + c;
    const int64_t m = (int64_t)e * N2 + n2;
    if (n2 >= N2 || m >= M) return make_double2(0.0, 0.0);
    const double* row = x + (offs ? offs[b] : (int64_t)b * ld) + m;
    double2 acc = make_double2(0.0, 0.0);
    for (int r = 0; r < step; ++r) {
      const double v = row[(int64_t)r * M];
      const double2 w = wr[r];
      acc.x = fma(v, w.x, acc.x);
      acc.y = fma(v, w.y, acc.y);
    }
    return cmul(acc, wm[m]);
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return;
    const size_t idx = (size_t)e * N2 + n2;
    ws[(size_t)b * N + idx] = cmul(v, tw[idx]);
  }
};

// searchsorted(lx, xo, 'right') - 1, clipped to [0, n_in-2]: the knot interval of an output abscissa (the same for every row)
__device__ __forceinline__ int64_t lerp_interval(const double* __restrict__ lx, int64_t n_in, double xo) {
  int64_t lo = 0, hi = n_in;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (lx[mid] <= xo) lo = mid + 1; else hi = mid;
  }
  return max((int64_t)0, min(lo - 1, n_in - 2));
}
__device__ __forceinline__ double lerp_eval(double y0, double y1, double x0, double x1, double xo) {
  return (y0 * (x1 - xo) + y1 * (xo - x0)) / (x1 - x0);
}

constexpr int LERP_ROWS = 16;  // rows per block: one interval search per output abscissa serves all of them

__global__ void __launch_bounds__(256) k_loglerp(const double* __restrict__ lx, const double* __restrict__ y, int64_t n_in,
                                                 int64_t ld_in, const double* __restrict__ lxo, int64_t n_out,
                                                 double* __restrict__ out, int64_t ld_out, int64_t batch) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  const double xo = lxo[j];
  const int64_t i = lerp_interval(lx, n_in, xo);
  const double x0 = lx[i], x1 = lx[i + 1];
  const int64_t b1 = min(batch, ((int64_t)blockIdx.y + 1) * LERP_ROWS);
  for (int64_t b = (int64_t)blockIdx.y * LERP_ROWS; b < b1; ++b) {
    const double* yy = y + (size_t)b * ld_in;
    out[(size_t)b * ld_out + j] = lerp_eval(yy[i], yy[i + 1], x0, x1, xo);
  }
}

// ---- gain grid of minimum_phase_impulse_response (autoeq/frequency_response.py:653-676) ----------------------------------
// dB curve on the log-f grid -> linear gain on linspace(0, fs//2, n): k=1 interpolation, the value at f_min for every
// f <= f_min (a prefix of n_below points), [normalize: - max - 0.5 dB], * 2, 10^(./20), last point 0.  Two launches:
// the interpolation (+ per-row maximum when normalising) and the shaping; arithmetic in the order the reference applies it.
__device__ __forceinline__ unsigned long long order_key(double v) {  // monotone map of doubles onto unsigned integers
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double order_value(unsigned long long k) {
  return __longlong_as_double((long long)((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k));
}
__global__ void __launch_bounds__(256) k_gain_lerp(const double* __restrict__ lx, const double* __restrict__ y, int64_t n_in,
                                                   int64_t ld_in, const double* __restrict__ lxo, int64_t n_out,
                                                   int64_t n_below, double lx_min, double* __restrict__ out, int64_t ld_out,
                                                   int64_t batch, unsigned long long* __restrict__ rowmax) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const bool live = j < n_out;
  const double xo = live ? (j < n_below ? lx_min : lxo[j]) : lx_min;
  const int64_t i = lerp_interval(lx, n_in, xo);
  const double x0 = lx[i], x1 = lx[i + 1];
  const int64_t b1 = min(batch, ((int64_t)blockIdx.y + 1) * LERP_ROWS);
  for (int64_t b = (int64_t)blockIdx.y * LERP_ROWS; b < b1; ++b) {
    const double* yy = y + (size_t)b * ld_in;
    const double v = lerp_eval(yy[i], yy[i + 1], x0, x1, xo);
    if (live) out[(size_t)b * ld_out + j] = v;
    if (rowmax) {
      unsigned long long k = live ? order_key(v) : 0ull;
      for (int o = 16; o; o >>= 1) k = max(k, __shfl_xor_sync(0xffffffffu, k, o));
      if ((threadIdx.x & 31) == 0) atomicMax(rowmax + b, k);
    }
  }
}
__global__ void __launch_bounds__(256) k_gain_shape(double* __restrict__ g, int64_t n, int64_t ld,
                                                    const unsigned long long* __restrict__ rowmax) {
  const int64_t b = blockIdx.y;
  double* row = g + (size_t)b * ld;
  const double mx = rowmax ? order_value(rowmax[b]) : 0.0;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    double v = row[j];
    if (rowmax) { v -= mx; v -= 0.5; }
    v *= 2.0;
    row[j] = j == n - 1 ? 0.0 : pow(10.0, v / 20.0);
  }
}

// ---- Savitzky-Golay (order 2, mode='interp') ------------------------------------------------------------
__device__ void savgol_pass(const double* __restrict__ src, double* __restrict__ dst, int n, int w,
                            const double* __restrict__ coef, const double* __restrict__ edge) {
  const int half = w / 2;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    if (i < half) {
      const double* e = edge + (size_t)i * w;
      for (int j = 0; j < w; ++j) acc += e[j] * src[j];
    } else if (i >= n - half) {
      const double* e = edge + (size_t)(n - 1 - i) * w;
      for (int j = 0; j < w; ++j) acc += e[j] * src[n - 1 - j];
    } else {
      // np.convolve(y, c, 'same')[i] = sum_j c[j] * y[i + half - j]
      for (int j = 0; j < w; ++j) acc += coef[j] * src[i + half - j];
    }
    dst[i] = acc;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) k_savgol_blend(const double* __restrict__ y, int n, int64_t ld, int wn, int itn,
                                                      const double* __restrict__ cn, const double* __restrict__ en, int wt,
                                                      int itt, const double* __restrict__ ct, const double* __restrict__ et,
                                                      const double* __restrict__ kt, double* __restrict__ out, int64_t ld_out) {
  extern __shared__ double sm[];
  double* a = sm;          // ping
  double* bb = sm + n;     // pong
  double* yn = sm + 2 * n; // result of the normal filter
  const int r = blockIdx.x;
  const double* src = y + (size_t)r * ld;
  for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = src[i];
  __syncthreads();
  double* cur = a;
  double* nxt = bb;
  for (int it = 0; it < itn; ++it) {
    savgol_pass(cur, nxt, n, wn, cn, en);
    double* t = cur; cur = nxt; nxt = t;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) yn[i] = cur[i];
  __syncthreads();
  if (itn > 0) {  // restart the treble chain from the input
    for (int i = threadIdx.x; i < n; i += blockDim.x) a[i] = src[i];
    __syncthreads();
  }
  cur = a;
  nxt = bb;
  for (int it = 0; it < itt; ++it) {
    savgol_pass(cur, nxt, n, wt, ct, et);
    double* t = cur; cur = nxt; nxt = t;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double k_t = kt[i];
    const double k_n = k_t * -1.0 + 1.0;
    out[(size_t)r * ld_out + i] = __dadd_rn(__dmul_rn(yn[i], k_n), __dmul_rn(cur[i], k_t));
  }
}

}  // namespace

// ================================================================================================
// minimum-phase FIR
// ================================================================================================
struct impb_minphase {
  impb_ctx* ctx = nullptr;
  int64_t n = 0;         // gain points == output taps
  int64_t numtaps = 0;   // 2n  (== n_fft of the homomorphic transform)
  int64_t nf = 0;        // irfft length of firwin2: 2 * 2^ceil(log2(numtaps))
  Fft2D<double> big;     // length nf
  Fft2D<double> small;   // length numtaps
  double* hamming = nullptr;  // [numtaps]
  double* nyq_S = nullptr;    // [nf/2 + 1] kernel of the Nyquist-bin imprint (see NyquistImprint below)
  double win_sum = 0.0;       // sum(hamming)
  void* tab = nullptr;        // FirwinTab[nf/2 + 1]: per-bin constants of the firwin2 synthesis for sampling rate tab_fs
  int tab_fs = 0;
  double2* pos_a = nullptr;   // [nf] position-order copies for the row-pass loader: {dx, df}
  double2* pos_b = nullptr;   //      {cos, sin} with the Hermitian sign / zeroing folded in
  int* pos_j = nullptr;       //      np.interp interval
  int64_t chunk = 1024;  // filters per pass over the chain (1.25 MB of workspace each at 9 600 taps): 896 filters take
                         // 2.98 / 2.19 / 1.94 ms in chunks of 64 / 256 / 896 -- launch and tail effects, not L2 residency
};

namespace {

// firwin2 spectrum evaluated at natural index k in [0, nf):  Hermitian extension of fx * shift.
//
// The homomorphic transform downstream takes log|H| at the Nyquist bin, where a type-II filter is
// analytically ZERO: what the reference feeds into the log there (~1e-12) is the imprint of the rounding
// of its phase ramp  -(numtaps-1)/2 * 1j * pi * x / nyq.  To land on the same value the ramp is rebuilt
// with the reference's exact sequence of IEEE operations (numpy divides a complex array by a real scalar
// by multiplying with fl(1/nyq); verified bit-identical on the host for every supported rate).
// Everything of the firwin2 grid that does not depend on the filter is tabulated once per (plan, sampling rate): where
// grid frequency x_k falls in the gain grid (np.interp's interval j, x - f_j, f_{j+1} - f_j), the phase-ramp factor
// exp(i a_k) with the reference's rounded angle a_k, and the Nyquist-imprint weight W_k = -(2/nf) d_k S_k.  The per-
// filter work of a bin is then two gain loads, one division and one multiply-add (np.interp's exact operation order)
// instead of an interval search and a large-argument sincos per (filter, bin): RowInvFirwin 270 -> see profiles/.
struct FirwinTab {
  double dx, df;   // x - f_j, f_{j+1} - f_j  (df = 0: take g[j] as is — last grid point or an exact hit)
  double cs, sn;   // cos / sin of fl(fl(b x) / nyq)
  double w;        // Nyquist-imprint weight
  int j, pad;
};

struct FirwinSpec {
  const double* gain;  // [batch, ld]
  int64_t ld, n, nf;
  double step_x;   // nyq / (nfreqs - 1)          : np.linspace(0, nyq, nfreqs)
  double nyq;
  double rnyq;     // fl(1 / nyq)
  double b;        // fl(-(numtaps - 1) / 2 * pi)
  double step_f;   // (fs // 2) / (n - 1)         : np.linspace(0, fs // 2, n)
  double f_stop;
  const FirwinTab* tab;
  const double2* pos_a;
  const double2* pos_b;
  const int* pos_j;
  __device__ __forceinline__ double fgrid(int64_t i) const { return i == n - 1 ? f_stop : __dmul_rn((double)i, step_f); }
  // grid frequency of firwin2 point kk in [0, nf/2], its np.interp interval j and the reference's phase-ramp angle
  __device__ __forceinline__ double xgrid(int64_t kk) const { return kk == nf / 2 ? nyq : __dmul_rn((double)kk, step_x); }
  __device__ __forceinline__ int64_t interval(double x) const {
    int64_t j = (int64_t)floor(x / step_f);
    j = max((int64_t)0, min(j, n - 1));
    while (j > 0 && fgrid(j) > x) --j;
    while (j < n - 1 && fgrid(j + 1) <= x) ++j;
    return j;
  }
  __device__ __forceinline__ double angle(double x) const { return __dmul_rn(__dmul_rn(b, x), rnyq); }
  // np.interp(x_kk, freq, gain) from the table
  __device__ __forceinline__ double fx_at(int bidx, const FirwinTab& t) const {
    const double* g = gain + (size_t)bidx * ld;
    const double g0 = g[t.j];
    if (t.df == 0.0) return g0;
    const double slope = __ddiv_rn(g[t.j + 1] - g0, t.df);
    return __dadd_rn(__dmul_rn(slope, t.dx), g0);
  }
  __device__ __forceinline__ double2 at(int bidx, int64_t k) const {
    const int64_t half = nf / 2;
    const bool mirror = k > half;
    const int64_t kk = mirror ? nf - k : k;
    const FirwinTab t = tab[kk];
    const double fx = fx_at(bidx, t);
    double2 v = make_double2(fx * t.cs, fx * t.sn);
    if (kk == 0 || kk == half) v.y = 0.0;  // c2r transforms ignore these imaginary parts
    if (mirror) v.y = -v.y;
    return v;
  }
  // the same value for the bin stored at workspace position `pos` (coalesced position-order tables)
  __device__ __forceinline__ double2 at_pos(int bidx, size_t pos) const {
    const double2 a = pos_a[pos], cssn = pos_b[pos];
    const double* g = gain + (size_t)bidx * ld;
    const int j = pos_j[pos];
    double fx = g[j];
    if (a.y != 0.0) fx = __dadd_rn(__dmul_rn(__ddiv_rn(g[j + 1] - fx, a.y), a.x), fx);
    return make_double2(fx * cssn.x, fx * cssn.y);
  }
};

// ---- Nyquist-bin imprint --------------------------------------------------------------------------------
// H[n_fft/2] = sum_n (-1)^n h[n] of the type-II filter is analytically zero; the ~1e-12 the reference's log sees
// there is the imprint of the ROUNDED phase-ramp angles a_k = fl(fl(b x_k) / nyq) of firwin2
// (scipy/signal/_fir_filter_design.py: fx2 = fx * exp(-(numtaps-1)/2 * 1j*pi*x/nyq)): with a_k = phi_k + d_k,
// phi_k = -pi (numtaps-1) k / nf the exact linear phase, bin k contributes (2/nf) fx_k cos(psi_kn + d_k) to h[n],
// psi_kn = pi k (2n - numtaps + 1) / nf.  The cos(psi) part is symmetric about the filter centre and cancels in
// the alternating sum; what is left is, to first order in d_k ~ 1e-12,
//     H[nyq] = -(2/nf) sum_k fx_k d_k S_k  +  (1/nf) fx_{nf/2} cos(a_{nf/2}) sum(w),
//     S_k = sum_n (-1)^n w[n] sin(psi_kn)                      (depends on the window only: plan constant).
// d_k is known to ~1e-28 (double-double pi, (numtaps-1) k / nf is exact because nf is a power of two), so this
// sum of 32 768 terms of size 1e-12 is good to ~1e-6 relative, whereas ANY double FFT — the reference's pocketfft
// included — only knows the bin to ~1e-16 absolute = 1e-4 relative.  The homomorphic transform turns a relative
// error e of this one bin into ~7e-5 e of the FIR peak (tools/minphase_localise.py), so a transform-computed value
// (1e-14 absolute with a different FFT algorithm) costs 5e-6 of peak while the analytic one lands within the
// reference's own noise (~1e-8 of peak).
__global__ void __launch_bounds__(256) k_mp_nyq_table(const double* __restrict__ win, int64_t numtaps, int64_t nf,
                                                      double* __restrict__ S) {
  const int64_t k = blockIdx.x;
  double acc = 0.0;
  for (int64_t n = threadIdx.x; n < numtaps; n += blockDim.x) {
    const int64_t m = (k * (2 * n - numtaps + 1)) % (2 * nf);  // psi = pi m / nf, reduced exactly in integers
    const double sn = sinpi((double)m / (double)nf);
    acc += (n & 1) ? -win[n] * sn : win[n] * sn;
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    S[k] = t;
  }
}

__global__ void __launch_bounds__(256) k_firwin_table(const FirwinSpec spec, int64_t numtaps, const double* __restrict__ S,
                                                      FirwinTab* __restrict__ tab) {
  const int64_t half = spec.nf / 2;
  const int64_t kk = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (kk > half) return;
  const double PI_HI = 3.141592653589793116, PI_LO = 1.2246467991473532e-16;
  const double x = spec.xgrid(kk);
  const int64_t j = spec.interval(x);
  FirwinTab t;
  t.j = (int)j;
  t.pad = 0;
  const double fj = spec.fgrid(j);
  if (j == spec.n - 1 || fj == x) { t.dx = 0.0; t.df = 0.0; }
  else { t.dx = x - fj; t.df = spec.fgrid(j + 1) - fj; }
  const double a = spec.angle(x);
  sincos(a, &t.sn, &t.cs);
  // d_k = a_k - phi_k with phi_k = -pi (numtaps - 1) k / nf in double-double ((numtaps-1) k / nf is exact)
  const double tt = (double)((numtaps - 1) * kk) / (double)spec.nf;
  const double p_hi = PI_HI * tt;
  const double p_lo = fma(PI_HI, tt, -p_hi) + PI_LO * tt;
  const double d = (a + p_hi) + p_lo;
  t.w = (kk >= 1 && kk < half) ? -(2.0 / (double)spec.nf) * d * S[kk] : 0.0;
  tab[kk] = t;
}

__global__ void __launch_bounds__(256) k_firwin_pos_table(const FirwinTab* __restrict__ tab, int64_t nf,
                                                          const int* __restrict__ perm1, const int* __restrict__ perm2,
                                                          int N1, int N2, double2* __restrict__ pos_a,
                                                          double2* __restrict__ pos_b, int* __restrict__ pos_j) {
  const int64_t pos = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (pos >= nf) return;
  const int row = (int)(pos / N2), e = (int)(pos % N2);
  const int64_t k = (int64_t)perm1[row] + (int64_t)N1 * perm2[e];
  const int64_t half = nf / 2;
  const bool mirror = k > half;
  const int64_t kk = mirror ? nf - k : k;
  const FirwinTab t = tab[kk];
  pos_a[pos] = make_double2(t.dx, t.df);
  pos_b[pos] = make_double2(t.cs, (kk == 0 || kk == half) ? 0.0 : (mirror ? -t.sn : t.sn));
  pos_j[pos] = t.j;
}

// H[nyq] of every filter: sum_k fx_k W_k (+ the firwin2 Nyquist term) as a fixed-order two-level sum (deterministic):
// grid (IMPRINT_SLICES, filters) partial sums, then one thread per filter adds them in slice order
constexpr int IMPRINT_SLICES = 16;
// np.interp of every filter's gain curve onto the firwin2 grid, once, in natural bin order (fully parallel and
// coalesced: one thread per (filter, bin)), fused with the partial sums of the Nyquist imprint.  The transform's loader
// then needs one 8-byte load per (filter, bin) instead of a dependent table -> interval -> gain chain at the low
// occupancy of a 98 KB-tile FFT pass.  fxn: [filters][nf/2 + 1].
__global__ void __launch_bounds__(256) k_firwin_fx(const FirwinSpec spec, double* __restrict__ fxn,
                                                   double* __restrict__ part) {
  const int bidx = blockIdx.y;
  const int64_t half = spec.nf / 2;
  const int64_t per = (half + 1 + IMPRINT_SLICES - 1) / IMPRINT_SLICES;
  const int64_t lo = blockIdx.x * per, hi = min(half + 1, (blockIdx.x + 1) * per);
  double acc = 0.0;
  for (int64_t k = lo + threadIdx.x; k < hi; k += blockDim.x) {
    const FirwinTab t = spec.tab[k];
    const double fx = spec.fx_at(bidx, t);
    fxn[(size_t)bidx * (half + 1) + k] = fx;
    acc += fx * t.w;  // t.w = 0 for k = 0 and k = nf/2
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    part[(size_t)bidx * IMPRINT_SLICES + blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) k_mp_nyq_partial(const FirwinSpec spec, double* __restrict__ part) {
  const int bidx = blockIdx.y;
  const int64_t half = spec.nf / 2;
  const int64_t per = (half + IMPRINT_SLICES - 1) / IMPRINT_SLICES;
  const int64_t lo = max((int64_t)1, blockIdx.x * per), hi = min(half, (blockIdx.x + 1) * per);
  double acc = 0.0;
#pragma unroll 4
  for (int64_t k = lo + threadIdx.x; k < hi; k += blockDim.x) {
    const FirwinTab t = spec.tab[k];
    acc += spec.fx_at(bidx, t) * t.w;
  }
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sm[w];
    part[(size_t)bidx * IMPRINT_SLICES + blockIdx.x] = t;
  }
}
__global__ void k_mp_nyq_finish(const FirwinSpec spec, const double* __restrict__ part, double win_sum, int rows,
                                double* __restrict__ out) {
  const int bidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (bidx >= rows) return;
  double t = 0.0;
  for (int q = 0; q < IMPRINT_SLICES; ++q) t += part[(size_t)bidx * IMPRINT_SLICES + q];
  const FirwinTab th = spec.tab[spec.nf / 2];
  out[bidx] = t + (spec.fx_at(bidx, th) * th.cs) * win_sum / (double)spec.nf;
}

struct RowInvFirwin {  // row DIT whose input is synthesised from the gain curve
  FirwinSpec spec;
  double2* ws;
  const double2* tw;
  int64_t N;
  int N1, N2;
  const int* perm1;
  const int* perm2;
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int row = t * TILE + c;
    if (row >= N1) return make_double2(0.0, 0.0);
    return spec.at(b, (int64_t)perm1[row] + (int64_t)N1 * perm2[e]);
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int row = t * TILE + c;
    if (row >= N1) return;
    const size_t idx = (size_t)row * N2 + e;
    ws[(size_t)b * N + idx] = cmul(v, cconj(tw[idx]));
  }
};

// column DIT -> real output row, scaled and windowed:  out[b, m] = Re(v) * scale * win[m]  (m < keep)
struct ColInvReal {
  const double2* ws;
  int64_t N;
  int N2;
  double* out;
  int64_t ld, keep;
  const double* win;  // may be nullptr
  double scale;
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return make_double2(0.0, 0.0);
    return ws[(size_t)b * N + (size_t)e * N2 + n2];
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int n2 = t * TILE + c;
    const int64_t m = (int64_t)e * N2 + n2;
    if (n2 >= N2 || m >= keep) return;
    double r = v.x * scale;
    if (win) r *= win[m];
    out[(size_t)b * ld + m] = r;
  }
};

// ---- two real filters per complex transform ---------------------------------------------------------------------------
// Every transform of the minimum-phase chain maps real rows to Hermitian spectra or back, so filters 2p and 2p+1 ride as
// re / im of ONE complex transform (the trick K1 plays with the two ears): half the butterflies, half the workspace
// traffic.  Inverse transforms pack trivially (Z = A + iB at every bin).  After a forward transform the two spectra are
// A[k] = (Z[k] + conj Z[N-k]) / 2 and B[k] = (Z[k] - conj Z[N-k]) / 2i; the loader of the NEXT (inverse) pass reads
// the mirror bin through the inverse digit-reversal maps, so no separate split pass exists.
struct Mirror {
  const int *perm1, *perm2, *inv1, *inv2;
  int N1, N2;
  int64_t N;
  // position of bin N - k for the bin at position (row, e); *k = its natural index
  __device__ __forceinline__ size_t of(int row, int e, int64_t* k) const {
    const int64_t kk = (int64_t)perm1[row] + (int64_t)N1 * perm2[e];
    *k = kk;
    const int64_t km = kk == 0 ? 0 : N - kk;
    return (size_t)inv1[(int)(km % N1)] * N2 + inv2[(int)(km / N1)];
  }
};
__device__ __forceinline__ void split_pair(double2 z, double2 zm, double2* a, double2* b) {
  *a = make_double2(0.5 * (z.x + zm.x), 0.5 * (z.y - zm.y));
  *b = make_double2(0.5 * (z.y + zm.y), -0.5 * (z.x - zm.x));
}

struct RowInvFirwin2 {  // row DIT over the synthesised spectra of filters 2b (re) and 2b+1 (im)
  FirwinSpec spec;
  int rows;  // filters in this launch
  double2* ws;
  const double2* tw;
  int64_t N;
  int N1, N2;
  const int* perm1;
  const int* perm2;
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int row = t * TILE + c;
    if (row >= N1) return make_double2(0.0, 0.0);
    const size_t pos = (size_t)row * N2 + e;
    const double2 a = spec.at_pos(2 * b, pos);
    if (2 * b + 1 >= rows) return a;
    const double2 bb = spec.at_pos(2 * b + 1, pos);
    return make_double2(a.x - bb.y, a.y + bb.x);
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int row = t * TILE + c;
    if (row >= N1) return;
    const size_t idx = (size_t)row * N2 + e;
    ws[(size_t)b * N + idx] = cmul(v, cconj(tw[idx]));
  }
};

// firwin2's irfft through the FORWARD engine: ifft(Z) = conj(fft(conj Z)) / N.  The synthesised spectrum is then read
// in NATURAL bin order by the column pass (coalesced table and gain-curve reads; as the input of the inverse engine
// it was gathered in digit-reversed order, four scattered 8-byte gain loads per bin, L2-bandwidth bound), and only
// the numtaps kept time samples (29 % of the 2^16) are scattered on the way out.
struct ColFwdFirwin2 {
  FirwinSpec spec;
  int rows;
  double2* ws;
  const double2* tw;
  int64_t N;
  int N2;
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return make_double2(0.0, 0.0);
    const int64_t k = (int64_t)e * N2 + n2;
    const double2 a = spec.at(2 * b, k);
    if (2 * b + 1 >= rows) return make_double2(a.x, -a.y);
    const double2 bb = spec.at(2 * b + 1, k);
    return make_double2(a.x - bb.y, -(a.y + bb.x));  // conj(A + iB)
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return;
    const size_t idx = (size_t)e * N2 + n2;
    ws[(size_t)b * N + idx] = cmul(v, tw[idx]);
  }
};

struct ColFwdFirwinFx2 {  // the same pass fed from the precomputed fx rows (k_firwin_fx)
  const double* fxn;        // [rows][nf/2 + 1]
  const FirwinTab* tab;
  int64_t nf;
  int rows;
  double2* ws;
  const double2* tw;
  int64_t N;
  int N2;
  __device__ __forceinline__ double2 one(int bidx, int64_t kk, bool mirror, const FirwinTab& t, int64_t half) const {
    const double fx = fxn[(size_t)bidx * (half + 1) + kk];
    double2 v = make_double2(fx * t.cs, fx * t.sn);
    if (kk == 0 || kk == half) v.y = 0.0;
    if (mirror) v.y = -v.y;
    return v;
  }
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return make_double2(0.0, 0.0);
    const int64_t k = (int64_t)e * N2 + n2, half = nf / 2;
    const bool mirror = k > half;
    const int64_t kk = mirror ? nf - k : k;
    const FirwinTab tb = tab[kk];
    const double2 a = one(2 * b, kk, mirror, tb, half);
    if (2 * b + 1 >= rows) return make_double2(a.x, -a.y);
    const double2 bb = one(2 * b + 1, kk, mirror, tb, half);
    return make_double2(a.x - bb.y, -(a.y + bb.x));  // conj(A + iB)
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return;
    const size_t idx = (size_t)e * N2 + n2;
    ws[(size_t)b * N + idx] = cmul(v, tw[idx]);
  }
};

struct RealOut2Op {  // conj(fft(conj Z))[m]: out[2b, m] = Re v, out[2b+1, m] = -Im v (scaled, windowed), m < keep
  double* out;
  int64_t ld, keep;
  int rows;
  const double* win;
  double scale;
  __device__ __forceinline__ void operator()(int b, size_t, int64_t m, double2 v) const {
    if (m >= keep) return;
    const double s = win ? scale * win[m] : scale;
    out[(size_t)(2 * b) * ld + m] = v.x * s;
    if (2 * b + 1 < rows) out[(size_t)(2 * b + 1) * ld + m] = -v.y * s;
  }
};

struct ColInvReal2 {  // column DIT -> two real rows: out[2b, m] = Re v, out[2b+1, m] = Im v (scaled, windowed)
  const double2* ws;
  int64_t N;
  int N2;
  double* out;
  int64_t ld, keep;
  int rows;
  const double* win;  // may be nullptr
  double scale;
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return make_double2(0.0, 0.0);
    return ws[(size_t)b * N + (size_t)e * N2 + n2];
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int n2 = t * TILE + c;
    const int64_t m = (int64_t)e * N2 + n2;
    if (n2 >= N2 || m >= keep) return;
    double r0 = v.x * scale, r1 = v.y * scale;
    if (win) { r0 *= win[m]; r1 *= win[m]; }
    out[(size_t)(2 * b) * ld + m] = r0;
    if (2 * b + 1 < rows) out[(size_t)(2 * b + 1) * ld + m] = r1;
  }
};

struct RealColIn2 {  // rows 2b (re) and 2b+1 (im) in, column DIF
  const double* x;
  int64_t ld, n;
  int rows;
  double2* ws;
  const double2* tw;
  int64_t N;
  int N2;
  const int64_t* offs = nullptr;  // optional start of every row (elements from x) instead of row * ld
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int n2 = t * TILE + c;
    const int64_t i = (int64_t)e * N2 + n2;
    if (n2 >= N2 || i >= n) return make_double2(0.0, 0.0);
    const bool two = 2 * b + 1 < rows;
    const int64_t o0 = offs ? offs[2 * b] : (int64_t)(2 * b) * ld;
    const int64_t o1 = two ? (offs ? offs[2 * b + 1] : (int64_t)(2 * b + 1) * ld) : 0;
    return make_double2(x[o0 + i], two ? x[o1 + i] : 0.0);
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int n2 = t * TILE + c;
    if (n2 >= N2) return;
    const size_t idx = (size_t)e * N2 + n2;
    ws[(size_t)b * N + idx] = cmul(v, tw[idx]);
  }
};

// min over k of the positive |A[k]|, |B[k]| of a paired spectrum (scipy: h_temp[h_temp > 0].min() per filter).
// grid = (slices, pairs); positive doubles order like their bit patterns, so the slices meet in an integer atomicMin
// (out must be preset to +inf).
__global__ void __launch_bounds__(256) k_min_positive_pair(const double2* __restrict__ ws, const Mirror mir,
                                                           const double* __restrict__ nyq_val, int rows,
                                                           double* __restrict__ out) {
  const int b = blockIdx.y;
  const double2* z = ws + (size_t)b * mir.N;
  double va = INFINITY, vb = INFINITY;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < mir.N; idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t k;
    const size_t im = mir.of((int)(idx / mir.N2), (int)(idx % mir.N2), &k);
    double2 a, bb;
    split_pair(z[idx], z[im], &a, &bb);
    double ma = hypot(a.x, a.y), mb = hypot(bb.x, bb.y);
    if (nyq_val && 2 * k == mir.N) {
      ma = fabs(nyq_val[2 * b]);
      if (2 * b + 1 < rows) mb = fabs(nyq_val[2 * b + 1]);
    }
    if (ma > 0.0) va = fmin(va, ma);
    if (mb > 0.0) vb = fmin(vb, mb);
  }
  for (int o = 16; o; o >>= 1) {
    va = fmin(va, __shfl_xor_sync(0xffffffffu, va, o));
    vb = fmin(vb, __shfl_xor_sync(0xffffffffu, vb, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(reinterpret_cast<unsigned long long*>(out + 2 * b), (unsigned long long)__double_as_longlong(va));
    if (2 * b + 1 < rows)
      atomicMin(reinterpret_cast<unsigned long long*>(out + 2 * b + 1), (unsigned long long)__double_as_longlong(vb));
  }
}

// K8, decimated: 20 log10 |X[first + step j]| of real rows 2b (re) and 2b+1 (im) from the position-ordered spectrum of
// the paired transform.  frequency_response() keeps every step-th bin only (core/impulse_response.py:170-178), so two
// rows share one complex transform and only the kept bins and their mirrors are read back.
__global__ void __launch_bounds__(256) k_pair_mag_db(const double2* __restrict__ ws, const Mirror mir, int rows,
                                                     int64_t first, int64_t step, int64_t n_out, double* __restrict__ out,
                                                     int64_t ld) {
  const int b = blockIdx.y;
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n_out) return;
  const int64_t k = first + step * j, km = k == 0 ? 0 : mir.N - k;
  const double2* z = ws + (size_t)b * mir.N;
  const double2 v = z[(size_t)mir.inv1[(int)(k % mir.N1)] * mir.N2 + mir.inv2[(int)(k / mir.N1)]];
  const double2 vm = z[(size_t)mir.inv1[(int)(km % mir.N1)] * mir.N2 + mir.inv2[(int)(km / mir.N1)]];
  double2 a, bb;
  split_pair(v, vm, &a, &bb);
  out[(size_t)(2 * b) * ld + j] = 20.0 * log10(hypot(a.x, a.y));
  if (2 * b + 1 < rows) out[(size_t)(2 * b + 1) * ld + j] = 20.0 * log10(hypot(bb.x, bb.y));
}

__global__ void k_fill_inf(double* __restrict__ p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = INFINITY;
}

struct RowInvPairLogMag {  // load 0.5 log(|A| + 1e-7 min+) + i 0.5 log(|B| + 1e-7 min+) from the paired spectrum
  const double2* src;
  const double* minpos;
  const double* nyq_val;
  int rows;
  Mirror mir;
  double2* ws;
  const double2* tw;
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int row = t * TILE + c;
    if (row >= mir.N1) return make_double2(0.0, 0.0);
    const double2* z = src + (size_t)b * mir.N;
    int64_t k;
    const size_t im = mir.of(row, e, &k);
    double2 a, bb;
    split_pair(z[(size_t)row * mir.N2 + e], z[im], &a, &bb);
    double ma = hypot(a.x, a.y), mb = hypot(bb.x, bb.y);
    const bool two = 2 * b + 1 < rows;
    if (nyq_val && 2 * k == mir.N) {
      ma = fabs(nyq_val[2 * b]);
      if (two) mb = fabs(nyq_val[2 * b + 1]);
    }
    const double la = 0.5 * log(ma + 1e-7 * minpos[2 * b]);
    const double lb = two ? 0.5 * log(mb + 1e-7 * minpos[2 * b + 1]) : 0.0;
    return make_double2(la, lb);
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int row = t * TILE + c;
    if (row >= mir.N1) return;
    const size_t idx = (size_t)row * mir.N2 + e;
    ws[(size_t)b * mir.N + idx] = cmul(v, cconj(tw[idx]));
  }
};

struct RowInvPairExp {  // load exp(A) + i exp(B) from the paired spectrum of two cepstra
  const double2* src;
  int rows;
  Mirror mir;
  double2* ws;
  const double2* tw;
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int row = t * TILE + c;
    if (row >= mir.N1) return make_double2(0.0, 0.0);
    const double2* z = src + (size_t)b * mir.N;
    int64_t k;
    const size_t im = mir.of(row, e, &k);
    double2 a, bb;
    split_pair(z[(size_t)row * mir.N2 + e], z[im], &a, &bb);
    double s, cs;
    sincos(a.y, &s, &cs);
    const double m = exp(a.x);
    double2 ea = make_double2(m * cs, m * s);
    if (2 * b + 1 >= rows) return ea;
    sincos(bb.y, &s, &cs);
    const double m2 = exp(bb.x);
    return make_double2(ea.x - m2 * s, ea.y + m2 * cs);
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int row = t * TILE + c;
    if (row >= mir.N1) return;
    const size_t idx = (size_t)row * mir.N2 + e;
    ws[(size_t)b * mir.N + idx] = cmul(v, cconj(tw[idx]));
  }
};

struct AbsOp {  // |X| in position order; the Nyquist bin takes the analytic imprint (k_mp_nyq_imprint) when given
  double* mag;
  int64_t N;
  const double* nyq_val;
  __device__ __forceinline__ void operator()(int b, size_t pos, int64_t k, double2 v) const {
    mag[(size_t)b * N + pos] = (nyq_val && 2 * k == N) ? fabs(nyq_val[b]) : hypot(v.x, v.y);
  }
};

struct ExpOp {  // exp(X) in position order, in place
  double2* ws;
  int64_t N;
  __device__ __forceinline__ void operator()(int b, size_t pos, int64_t, double2 v) const {
    double s, c;
    sincos(v.y, &s, &c);
    const double m = exp(v.x);
    ws[(size_t)b * N + pos] = make_double2(m * c, m * s);
  }
};

__global__ void __launch_bounds__(256) k_min_positive(const double* __restrict__ mag, int64_t N, double* __restrict__ out) {
  const int b = blockIdx.x;
  const double* m = mag + (size_t)b * N;
  double v = INFINITY;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    const double t = m[i];
    if (t > 0.0) v = fmin(v, t);
  }
  for (int o = 16; o; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) v = fmin(v, sm[w]);
    out[b] = v;
  }
}

struct RowInvLogMag {  // load 0.5*log(|X| + 1e-7*minpos) as a real spectrum (position order)
  const double* mag;
  const double* minpos;
  double2* ws;
  const double2* tw;
  int64_t N;
  int N1, N2;
  __device__ __forceinline__ double2 load(int b, int t, int e, int c) const {
    const int row = t * TILE + c;
    if (row >= N1) return make_double2(0.0, 0.0);
    const double v = mag[(size_t)b * N + (size_t)row * N2 + e] + 1e-7 * minpos[b];
    return make_double2(0.5 * log(v), 0.0);
  }
  __device__ __forceinline__ void store(int b, int t, int e, int c, double2 v) const {
    const int row = t * TILE + c;
    if (row >= N1) return;
    const size_t idx = (size_t)row * N2 + e;
    ws[(size_t)b * N + idx] = cmul(v, cconj(tw[idx]));
  }
};

// homomorphic window  lmin[n] = 2u[n] - d[n]
__global__ void k_homomorphic_window(double* __restrict__ w, int64_t n_fft) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n_fft) return;
  const int64_t stop = n_fft / 2;
  double v = 0.0;
  if (i == 0) v = 1.0;
  else if (i < stop) v = 2.0;
  else if (i == stop) v = 1.0 + (double)(n_fft % 2);
  w[i] = v;
}


// ================================================================================================
// K21 — FrequencyResponse.equalize  (autoeq/frequency_response.py:1241-1310)
// ================================================================================================
// One CTA per curve, everything in shared memory (F <= EQ_MAXF points):
//   gain = -error * gain_k;  clipped = gain > cap;  eq = where(clipped, cap, gain)
//   kinks = indices i >= 1 where clipped changes;  every point within `half` of a kink is dropped, except the last two
//   eq = InterpolatedUnivariateSpline(logf[keep], eq[keep], k=2)(logf)
// The k = 2 interpolating spline of FITPACK (fpcurf with s = 0) puts its interior knots midway between the data sites,
// t = [x0, x0, x0, (x1+x2)/2, ..., (x_{m-3}+x_{m-2})/2, x_{m-1}, x_{m-1}, x_{m-1}], so every site x_i (0 < i < m-1)
// lies in knot interval i+1 and the collocation matrix is TRIDIAGONAL (B_{i-1}, B_i, B_{i+1} at x_i; c_0 = y_0,
// c_{m-1} = y_{m-1}).  It is totally positive, so elimination without pivoting is stable; FITPACK solves the same
// system with Givens rotations — agreement is to rounding (tests: <= 1e-9 dB against the reference golden).
constexpr int EQ_MAXF = 2048;

// Shared-memory layout of one curve, sized by its length (783 points: 47 KB, four curves per SM): seven double
// arrays of P = n + 4 entries and the clip flags.
struct EqShared {
  double *x, *y;   // kept sites (log10 f) / kept values, then spline coefficients
  double *eq;      // clipped gain at every point; after the compaction: B-spline value h0 of every collocation row
  double *t;       // knots
  double *cp;      // keep flags, then the modified upper diagonal of the Thomas elimination
  double *h1, *h2;
  int* keep;
  int* wsum;       // per-warp counts of the ordered compaction
  __device__ EqShared(unsigned char* raw, int n) {
    const int P = (n + 4) & ~1;
    double* d = reinterpret_cast<double*>(raw);
    x = d; y = d + P; eq = d + 2 * P; t = d + 3 * P; cp = d + 4 * P; h1 = d + 5 * P; h2 = d + 6 * P;
    keep = reinterpret_cast<int*>(d + 7 * P);
    wsum = keep + P;
  }
  static size_t bytes(int n) { const size_t P = (n + 4) & ~1; return 7 * P * sizeof(double) + (P + 8) * sizeof(int); }
};

// quadratic B-spline values N_{l-2}, N_{l-1}, N_l at x inside knot interval [t_l, t_{l+1})  (de Boor recursion, fpbspl)
__device__ __forceinline__ void bspl2(const double* t, int l, double x, double* h) {
  double h0 = 1.0, h1, h2;
  {  // degree 1
    const double f = h0 / (t[l + 1] - t[l]);
    h0 = f * (t[l + 1] - x);
    h1 = f * (x - t[l]);
  }
  {  // degree 2
    const double f0 = h0 / (t[l + 1] - t[l - 1]);
    const double f1 = h1 / (t[l + 2] - t[l]);
    h2 = f1 * (x - t[l]);
    h1 = f0 * (x - t[l - 1]) + f1 * (t[l + 2] - x);
    h0 = f0 * (t[l + 1] - x);
  }
  h[0] = h0; h[1] = h1; h[2] = h2;
}

__global__ void __launch_bounds__(256) k_equalize(const double* __restrict__ error, int64_t ld_err,
                                                  const double* __restrict__ logf, const double* __restrict__ cap,
                                                  const double* __restrict__ gain_k, int n, int half, int smoothen,
                                                  double* __restrict__ out, int64_t ld_out, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char eq_raw[];
  const EqShared sh(eq_raw, n);
  const int b = blockIdx.x;
  const double* err = error + (size_t)b * ld_err;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double g = -err[i] * gain_k[i];
    const bool cl = g > cap[i];
    sh.eq[i] = cl ? cap[i] : g;
    sh.keep[i] = cl ? 3 : 1;  // bit 1: clipped flag (scratch), bit 0: keep
  }
  __syncthreads();
  if (!smoothen) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[(size_t)b * ld_out + i] = sh.eq[i];
    if (threadIdx.x == 0 && status) status[b] = 0;
    return;
  }
  // kink windows: a point is dropped when a change of the clipped flag lies within `half` points of it
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    bool doomed = false;
    const int lo = max(i - half, 1), hi = min(i + half, n - 1);
    for (int k = lo; k <= hi && !doomed; ++k) doomed = ((sh.keep[k] >> 1) != (sh.keep[k - 1] >> 1));
    if (i >= n - 2) doomed = false;  // the last two points are never dropped
    sh.cp[i] = doomed ? 0.0 : 1.0;
  }
  __syncthreads();
  // ordered compaction of the kept points, 256 at a time (ballot ranks inside a warp, warp counts through shared memory)
  int m = 0;
  for (int i0 = 0; i0 < n; i0 += 256) {
    const int i = i0 + threadIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool kept = i < n && sh.cp[i] != 0.0;
    const unsigned bal = __ballot_sync(0xffffffffu, kept);
    if (lane == 0) sh.wsum[w] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int cnt = sh.wsum[q];
      before += q < w ? cnt : 0;
      total += cnt;
    }
    if (kept) {
      const int pos = m + before + __popc(bal & ((1u << lane) - 1u));
      sh.x[pos] = logf[i];
      sh.y[pos] = sh.eq[i];
    }
    m += total;
    __syncthreads();
  }
  if (m < 3) {  // dfitpack raises for m <= k; the host turns the status into the same kind of error
    if (threadIdx.x == 0 && status) status[b] = 1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[(size_t)b * ld_out + i] = sh.eq[i];
    return;
  }
  for (int j = threadIdx.x; j < m + 3; j += blockDim.x) {
    double v;
    if (j < 3) v = sh.x[0];
    else if (j >= m) v = sh.x[m - 1];
    else v = 0.5 * (sh.x[j - 2] + sh.x[j - 1]);  // t[3 + i] = (x[i+1] + x[i+2]) / 2
    sh.t[j] = v;
  }
  __syncthreads();
  // the three B-spline values of every collocation row, by all threads (eq is free after the compaction): the serial
  // elimination below is then one multiply-add and two divisions per row instead of seven divisions
  double* h0 = sh.eq;
  for (int i = 1 + threadIdx.x; i <= m - 2; i += blockDim.x) {
    double h[3];
    bspl2(sh.t, i + 1, sh.x[i], h);
    h0[i] = h[0]; sh.h1[i] = h[1]; sh.h2[i] = h[2];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // tridiagonal collocation system: row i (1 <= i <= m-2): a_i c_{i-1} + d_i c_i + u_i c_{i+1} = y_i
    double* c = sh.y;
    double* cpp = sh.cp;  // modified upper diagonal
    cpp[0] = 0.0;         // row 0: c_0 = y_0
    for (int i = 1; i <= m - 2; ++i) {
      const double a = h0[i];
      const double den = sh.h1[i] - a * cpp[i - 1];
      cpp[i] = sh.h2[i] / den;
      c[i] = (c[i] - a * c[i - 1]) / den;
    }
    // row m-1: c_{m-1} = y_{m-1}; back substitution
    for (int i = m - 2; i >= 1; --i) c[i] -= cpp[i] * c[i + 1];
  }
  __syncthreads();
  // evaluate at every grid point (splev): interval by binary search over the interior knots t[3 .. m-1]
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double xv = fmin(fmax(logf[i], sh.t[2]), sh.t[m]);
    int lo = 2, hi = m - 1;  // largest l in [2, m-1] with t[l] <= xv
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (sh.t[mid] <= xv) lo = mid; else hi = mid - 1;
    }
    double h[3];
    bspl2(sh.t, lo, xv, h);
    out[(size_t)b * ld_out + i] = h[0] * sh.y[lo - 2] + h[1] * sh.y[lo - 1] + h[2] * sh.y[lo];
  }
  if (threadIdx.x == 0 && status) status[b] = 0;
}

}  // namespace

extern "C" {

int32_t impb_fft_create(impb_ctx* ctx, int64_t n, int32_t precision, impb_fft** out) {
  if (!ctx || !out || n < 4 || precision != 64) {
    set_error("impb_fft_create: bad argument (n=%lld, precision=%d; only 64 is implemented)", (long long)n, precision);
    return IMPB_EINVAL;
  }
  DeviceGuard g(ctx->device);
  Plan2D p;
  if (plan_2d(ctx, n, true, 8, &p)) {
    impb_fft* f = new impb_fft();
    f->ctx = ctx;
    f->n = n;
    int rc = f->f.init(p);
    if (rc) { delete f; return rc; }
    *out = f;
    return IMPB_OK;
  }
  // Bluestein over a 5-smooth circular length M >= 2n - 1
  if (!plan_2d(ctx, 2 * n - 1, false, 8, &p)) {
    set_error("impb_fft_create: no plan for length %lld (direct or chirp-z)", (long long)n);
    return IMPB_EUNSUPPORTED;
  }
  impb_fft* f = new impb_fft();
  f->ctx = ctx;
  f->n = n;
  f->chirp = true;
  f->M = p.N;
  int rc = f->f.init(p);
  if (rc) { delete f; return rc; }
  const int64_t M = p.N;
  std::vector<double2> w((size_t)n), h((size_t)M, make_double2(0.0, 0.0));
  for (int64_t m = 0; m < n; ++m) {  // phase pi m^2 / n with m^2 reduced modulo 2n in integers
    const unsigned __int128 sq = (unsigned __int128)(uint64_t)m * (uint64_t)m;
    const int64_t r = (int64_t)(sq % (unsigned __int128)(2 * n));
    const long double ang = 3.141592653589793238462643383279502884L * (long double)r / (long double)n;
    w[(size_t)m] = make_double2((double)cosl(ang), (double)sinl(ang));
    h[(size_t)m] = w[(size_t)m];
    if (m) h[(size_t)(M - m)] = w[(size_t)m];
  }
  double2* d_h = nullptr;
  cudaError_t e = cudaMalloc(&f->w, sizeof(double2) * n);
  if (e == cudaSuccess) e = cudaMalloc(&f->Hs, sizeof(double2) * M);
  if (e == cudaSuccess) e = cudaMalloc(&d_h, sizeof(double2) * M);
  if (e == cudaSuccess) e = cudaMemcpy(f->w, w.data(), sizeof(double2) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_h, h.data(), sizeof(double2) * M, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(d_h); impb_fft_destroy(f); return cuda_fail(e, "chirp tables"); }
  ColPlain<double, TILE> ci{d_h, f->Hs, f->f.tw, M, p.N2, 0};
  rc = f->f.col_fwd<TILE>(ci, 1, 0);
  if (!rc) rc = f->f.row_fwd<TILE>(RowPlain<double, TILE>{f->Hs, f->Hs, nullptr, M, p.N1, p.N2, 0, 0}, 1, 0);
  e = cudaDeviceSynchronize();
  cudaFree(d_h);
  if (rc || e != cudaSuccess) { impb_fft_destroy(f); return rc ? rc : cuda_fail(e, "chirp spectrum"); }
  *out = f;
  return IMPB_OK;
}

void impb_fft_destroy(impb_fft* f) {
  if (!f) return;
  DeviceGuard g(f->ctx->device);
  f->f.release();
  cudaFree(f->w);
  cudaFree(f->Hs);
  for (auto& kv : f->fold_w) cudaFree(kv.second);
  delete f;
}

int64_t impb_fft_workspace_bytes(const impb_fft* f, int64_t batch) {
  if (!f || batch < 1) return 0;
  return batch * (f->chirp ? f->M : f->n) * (int64_t)sizeof(double2);
}

int32_t impb_magnitude_db(impb_fft* f, const double* x, int64_t ld_x, double* mag_out, int64_t ld_out, int64_t batch,
                          void* workspace, void* stream) {
  if (!f || !x || !mag_out || !workspace || batch < 0 || ld_x < f->n) {
    set_error("impb_magnitude_db: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0) return IMPB_OK;
  DeviceGuard g(f->ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const Plan2D& p = f->f.p;
  const int64_t half = (f->n + 1) / 2;
  if (ld_out < half) { set_error("impb_magnitude_db: ld_out < ceil(n/2)"); return IMPB_EINVAL; }
  double2* ws = (double2*)workspace;
  if (f->chirp) {
    for (int64_t b0 = 0; b0 < batch; b0 += 32768) {
      const int nb = (int)std::min<int64_t>(32768, batch - b0);
      double2* w = ws + (size_t)b0 * p.N;
      int rc = f->f.col_fwd<TILE>(ChirpColIn{x + (size_t)b0 * ld_x, ld_x, f->n, f->w, w, f->f.tw, p.N, p.N2}, nb, st);
      if (!rc) rc = f->f.row_fwd<TILE>(RowPlain<double, TILE>{w, w, nullptr, p.N, p.N1, p.N2, 0, 0}, nb, st);
      if (rc) return rc;
      k_cmul_rows<<<dim3((unsigned)std::min<int64_t>((p.N + 255) / 256, 2048), nb), 256, 0, st>>>(w, f->Hs, p.N);
      rc = f->f.row_inv<TILE>(RowPlain<double, TILE>{w, w, f->f.tw, p.N, p.N1, p.N2, 1, 0}, nb, st);
      if (!rc) rc = f->f.col_inv<TILE>(ChirpColOut{w, p.N, p.N2, mag_out + (size_t)b0 * ld_out, ld_out, half,
                                                   1.0 / (double)p.N}, nb, st);
      if (rc) return rc;
    }
    IMPB_CUDA(cudaGetLastError());
    return IMPB_OK;
  }
  for (int64_t b0 = 0; b0 < batch; b0 += 32768) {
    const int nb = (int)std::min<int64_t>(32768, batch - b0);
    RealColIn ci{x + (size_t)b0 * ld_x, ld_x, f->n, ws + (size_t)b0 * p.N, f->f.tw, p.N, p.N2};
    int rc = f->f.col_fwd<TILE>(ci, nb, st);
    if (rc) return rc;
    RowFwdOp<MagDbOp> ro{ws + (size_t)b0 * p.N, p.N, p.N1, p.N2, f->f.perm_col, f->f.perm_row,
                         MagDbOp{mag_out + (size_t)b0 * ld_out, ld_out, half}};
    if ((rc = f->f.row_fwd<TILE>(ro, nb, st))) return rc;
  }
  return IMPB_OK;
}

int32_t impb_magnitude_db_decimated(impb_fft* f, const double* x, int64_t ld_x, const int64_t* row_offsets, int64_t first,
                                    int64_t step, double* mag_out, int64_t ld_out, int64_t batch, void* workspace,
                                    void* stream) {
  if (!f || !x || !mag_out || !workspace || batch < 0 || (!row_offsets && ld_x < f->n) || first < 0 || step < 1) {
    set_error("impb_magnitude_db_decimated: bad argument");
    return IMPB_EINVAL;
  }
  if (f->chirp) {
    set_error("impb_magnitude_db_decimated: chirp-z plans only serve impb_magnitude_db");
    return IMPB_EUNSUPPORTED;
  }
  const int64_t half = (f->n + 1) / 2;
  const int64_t n_out = half > first ? (half - first + step - 1) / step : 0;
  if (ld_out < n_out) { set_error("impb_magnitude_db_decimated: ld_out < kept bins"); return IMPB_EINVAL; }
  if (batch == 0 || n_out == 0) return IMPB_OK;
  DeviceGuard g(f->ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const Plan2D& p = f->f.p;
  const Mirror mir{f->f.perm_col, f->f.perm_row, f->f.inv_col, f->f.inv_row, p.N1, p.N2, p.N};
  double2* ws = (double2*)workspace;
  for (int64_t b0 = 0; b0 < batch; b0 += 65534) {  // an even number of rows per launch: pairs never straddle
    const int nr = (int)std::min<int64_t>(65534, batch - b0), np = (nr + 1) / 2;
    double2* w = ws + (size_t)(b0 / 2) * p.N;
    const RealColIn2 ci{row_offsets ? x : x + (size_t)b0 * ld_x, ld_x, f->n, nr, w, f->f.tw, p.N, p.N2,
                        row_offsets ? row_offsets + b0 : nullptr};
    int rc = f->f.col_fwd<TILE>(ci, np, st);
    if (!rc) rc = f->f.row_fwd<TILE>(RowPlain<double, TILE>{w, w, nullptr, p.N, p.N1, p.N2, 0, 0}, np, st);
    if (rc) return rc;
    k_pair_mag_db<<<dim3((unsigned)((n_out + 255) / 256), np), 256, 0, st>>>(w, mir, nr, first, step, n_out,
                                                                              mag_out + (size_t)b0 * ld_out, ld_out);
  }
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

int32_t impb_magnitude_db_folded(impb_fft* f, const double* x, int64_t ld_x, const int64_t* row_offsets, int64_t first,
                                 int64_t step, double* mag_out, int64_t ld_out, int64_t batch, void* workspace,
                                 void* stream) {
  if (!f || !x || !mag_out || !workspace || batch < 0 || first < 0 || step < 2 || step > 64 || first >= step ||
      (!row_offsets && ld_x < f->n * step)) {
    set_error("impb_magnitude_db_folded: bad argument");
    return IMPB_EINVAL;
  }
  if (f->chirp || !(f->f.col_fixed && f->f.row_fixed)) {
    // a chirp-z plan, or sub-FFT lengths outside the specialised menu (fft2d.cuh): the generic engine is slower on the
    // folded row than the specialised one on the paired full row (57 600 / 5: 0.92 vs 0.53 ms per 896 rows)
    set_error("impb_magnitude_db_folded: no specialised plan for the folded length; use impb_magnitude_db_decimated");
    return IMPB_EUNSUPPORTED;
  }
  const int64_t M = f->n, n_full = M * step;
  const int64_t half = (n_full + 1) / 2;
  const int64_t n_out = half > first ? (half - first + step - 1) / step : 0;  // <= M
  if (ld_out < n_out) { set_error("impb_magnitude_db_folded: ld_out < kept bins"); return IMPB_EINVAL; }
  if (batch == 0 || n_out == 0) return IMPB_OK;
  DeviceGuard g(f->ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  double2* fold_w = nullptr;
  {
    std::lock_guard<std::mutex> lk(f->fold_mu);
    auto it = f->fold_w.find({first, step});
    if (it == f->fold_w.end()) {
      std::vector<double2> w((size_t)(M + step));
      const long double tau = -2.0L * 3.141592653589793238462643383279502884L;
      for (int64_t m = 0; m < M; ++m) {  // first * m < step * M = n_full: the angle stays inside one turn
        const long double a = tau * (long double)(first * m) / (long double)n_full;
        w[(size_t)m] = make_double2((double)cosl(a), (double)sinl(a));
      }
      for (int64_t r = 0; r < step; ++r) {
        const long double a = tau * (long double)((first * r) % step) / (long double)step;
        w[(size_t)(M + r)] = make_double2((double)cosl(a), (double)sinl(a));
      }
      double2* d = nullptr;
      IMPB_CUDA(cudaMalloc(&d, sizeof(double2) * (size_t)(M + step)));
      cudaError_t e = cudaMemcpy(d, w.data(), sizeof(double2) * (size_t)(M + step), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) { cudaFree(d); return cuda_fail(e, "fold tables"); }
      it = f->fold_w.emplace(std::make_pair(first, step), d).first;
    }
    fold_w = it->second;
  }
  const Plan2D& p = f->f.p;
  double2* ws = (double2*)workspace;
  for (int64_t b0 = 0; b0 < batch; b0 += 32768) {
    const int nb = (int)std::min<int64_t>(32768, batch - b0);
    double2* w = ws + (size_t)b0 * p.N;
    const FoldColIn ci{row_offsets ? x : x + (size_t)b0 * ld_x, ld_x, row_offsets ? row_offsets + b0 : nullptr, M, (int)step,
                       fold_w, fold_w + M, w, f->f.tw, p.N, p.N2};
    int rc = f->f.col_fwd<TILE>(ci, nb, st);
    if (rc) return rc;
    RowFwdOp<MagDbOp> ro{w, p.N, p.N1, p.N2, f->f.perm_col, f->f.perm_row,
                         MagDbOp{mag_out + (size_t)b0 * ld_out, ld_out, n_out}};
    if ((rc = f->f.row_fwd<TILE>(ro, nb, st))) return rc;
  }
  return IMPB_OK;
}

int32_t impb_magnitude_lin(impb_fft* f, const double* x, int64_t ld_x, int64_t n_valid, double* mag_out, int64_t ld_out,
                           int64_t n_bins, int64_t batch, void* workspace, void* stream) {
  if (!f || !x || !mag_out || !workspace || batch < 0 || n_valid < 0 || n_valid > f->n || ld_x < n_valid || n_bins < 0 ||
      n_bins > f->n / 2 + 1 || ld_out < n_bins) {
    set_error("impb_magnitude_lin: bad argument");
    return IMPB_EINVAL;
  }
  if (f->chirp) { set_error("impb_magnitude_lin: chirp-z plans only serve impb_magnitude_db"); return IMPB_EUNSUPPORTED; }
  if (batch == 0 || n_bins == 0) return IMPB_OK;
  DeviceGuard g(f->ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const Plan2D& p = f->f.p;
  double2* ws = (double2*)workspace;
  for (int64_t b0 = 0; b0 < batch; b0 += 32768) {
    const int nb = (int)std::min<int64_t>(32768, batch - b0);
    RealColIn ci{x + (size_t)b0 * ld_x, ld_x, n_valid, ws + (size_t)b0 * p.N, f->f.tw, p.N, p.N2};
    int rc = f->f.col_fwd<TILE>(ci, nb, st);
    if (rc) return rc;
    RowFwdOp<MagLinOp> ro{ws + (size_t)b0 * p.N, p.N, p.N1, p.N2, f->f.perm_col, f->f.perm_row,
                          MagLinOp{mag_out + (size_t)b0 * ld_out, ld_out, n_bins}};
    if ((rc = f->f.row_fwd<TILE>(ro, nb, st))) return rc;
  }
  return IMPB_OK;
}

int32_t impb_loglerp(impb_ctx* ctx, const double* logf_in, const double* y_in, int64_t n_in, int64_t ld_in,
                     const double* logf_out, int64_t n_out, double* y_out, int64_t ld_out, int64_t batch, void* stream) {
  if (!ctx || !logf_in || !y_in || !logf_out || !y_out || n_in < 2 || n_out < 0 || batch < 0 || ld_in < n_in ||
      ld_out < n_out) {
    set_error("impb_loglerp: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0 || n_out == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  const int thr = 256;
  const unsigned bx = (unsigned)((n_out + thr - 1) / thr);
  const int64_t chunk = 65535 * (int64_t)LERP_ROWS;
  for (int64_t b0 = 0; b0 < batch; b0 += chunk) {
    const int64_t nb = std::min<int64_t>(chunk, batch - b0);
    k_loglerp<<<dim3(bx, (unsigned)((nb + LERP_ROWS - 1) / LERP_ROWS)), thr, 0, (cudaStream_t)stream>>>(
        logf_in, y_in + (size_t)b0 * ld_in, n_in, ld_in, logf_out, n_out, y_out + (size_t)b0 * ld_out, ld_out, nb);
  }
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

int32_t impb_minphase_gain_grid(impb_ctx* ctx, const double* logf_in, const double* eq_db, int64_t n_in, int64_t ld_in,
                                const double* logf_out, int64_t n_out, int64_t n_below, double logf_min, int32_t normalize,
                                double* gain_out, int64_t ld_out, int64_t batch, void* row_scratch, void* stream) {
  if (!ctx || !logf_in || !eq_db || !logf_out || !gain_out || n_in < 2 || n_out < 1 || batch < 0 || ld_in < n_in ||
      ld_out < n_out || n_below < 0 || n_below > n_out || (normalize && !row_scratch) || batch > 65535 * (int64_t)LERP_ROWS) {
    set_error("impb_minphase_gain_grid: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* rowmax = normalize ? (unsigned long long*)row_scratch : nullptr;
  if (rowmax) IMPB_CUDA(cudaMemsetAsync(rowmax, 0, sizeof(unsigned long long) * batch, st));  // key 0 < key(any double)
  const unsigned bx = (unsigned)((n_out + 255) / 256);
  k_gain_lerp<<<dim3(bx, (unsigned)((batch + LERP_ROWS - 1) / LERP_ROWS)), 256, 0, st>>>(
      logf_in, eq_db, n_in, ld_in, logf_out, n_out, n_below, logf_min, gain_out, ld_out, batch, rowmax);
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = (int)std::min<int64_t>(65535, batch - b0);
    k_gain_shape<<<dim3(std::min(bx, 16u), nb), 256, 0, st>>>(gain_out + (size_t)b0 * ld_out, n_out, ld_out,
                                                              rowmax ? rowmax + b0 : nullptr);
  }
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

int32_t impb_savgol_blend(impb_ctx* ctx, const double* y, int64_t n, int64_t ld, int32_t w_normal, int32_t it_normal,
                          const double* coef_normal, const double* edge_normal, int32_t w_treble, int32_t it_treble,
                          const double* coef_treble, const double* edge_treble, const double* k_treble, double* out,
                          int64_t ld_out, int64_t batch, void* stream) {
  if (!ctx || !y || !out || !k_treble || !coef_normal || !edge_normal || !coef_treble || !edge_treble || n < 1 ||
      ld < n || ld_out < n || batch < 0 || w_normal < 3 || w_treble < 3 || !(w_normal & 1) || !(w_treble & 1) ||
      it_normal < 0 || it_treble < 0) {
    set_error("impb_savgol_blend: bad argument");
    return IMPB_EINVAL;
  }
  if (w_normal > n || w_treble > n) {
    set_error("If mode is 'interp', window_length must be less than or equal to the size of x.");
    return IMPB_EINVAL;
  }
  if (batch == 0) return IMPB_OK;
  const size_t smem = 3 * (size_t)n * sizeof(double);
  if (smem > ctx->smem_optin) {
    set_error("impb_savgol_blend: curve of %lld points does not fit shared memory", (long long)n);
    return IMPB_EUNSUPPORTED;
  }
  DeviceGuard g(ctx->device);
  IMPB_CUDA(cudaFuncSetAttribute(k_savgol_blend, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_savgol_blend<<<(unsigned)batch, 256, smem, (cudaStream_t)stream>>>(y, (int)n, ld, w_normal, it_normal, coef_normal,
                                                                      edge_normal, w_treble, it_treble, coef_treble,
                                                                      edge_treble, k_treble, out, ld_out);
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

// ---- minimum phase ------------------------------------------------------------------------------------
int32_t impb_minphase_create(impb_ctx* ctx, int64_t n, const double* hamming_host, impb_minphase** out) {
  if (!ctx || !out || n < 4 || !hamming_host) {
    set_error("impb_minphase_create: bad argument");
    return IMPB_EINVAL;
  }
  DeviceGuard g(ctx->device);
  impb_minphase* mp = new impb_minphase();
  if (const char* e = getenv("IMPB_MINPHASE_CHUNK")) {  // A/B: filters per chunk (even)
    const long v = atol(e);
    if (v >= 2) mp->chunk = v & ~1L;
  }
  mp->ctx = ctx;
  mp->n = n;
  mp->numtaps = 2 * n;
  int64_t pw = 1;
  while (pw < mp->numtaps) pw <<= 1;
  mp->nf = 2 * pw;  // nfreqs = 1 + 2^ceil(log2(numtaps));  irfft length = 2 (nfreqs - 1)
  Plan2D pb, ps;
  if (!plan_2d(ctx, mp->nf, true, 8, &pb) || !plan_2d(ctx, mp->numtaps, true, 8, &ps)) {
    set_error("impb_minphase_create: no FFT plan for n=%lld (2n must be 5-smooth)", (long long)n);
    delete mp;
    return IMPB_EUNSUPPORTED;
  }
  int rc = mp->big.init(pb);
  if (!rc) rc = mp->small.init(ps);
  if (!rc) {
    cudaError_t e = cudaMalloc(&mp->hamming, mp->numtaps * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(mp->hamming, hamming_host, mp->numtaps * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&mp->nyq_S, (mp->nf / 2 + 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&mp->tab, (mp->nf / 2 + 1) * sizeof(FirwinTab));
    if (e == cudaSuccess) e = cudaMalloc(&mp->pos_a, mp->nf * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc(&mp->pos_b, mp->nf * sizeof(double2));
    if (e == cudaSuccess) e = cudaMalloc(&mp->pos_j, mp->nf * sizeof(int));
    if (e == cudaSuccess) {
      k_mp_nyq_table<<<(unsigned)(mp->nf / 2 + 1), 256>>>(mp->hamming, mp->numtaps, mp->nf, mp->nyq_S);
      e = cudaDeviceSynchronize();
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "hamming upload / Nyquist imprint table");
    for (int64_t i = 0; i < mp->numtaps; ++i) mp->win_sum += hamming_host[i];
  }
  if (rc) { impb_minphase_destroy(mp); return rc; }
  *out = mp;
  return IMPB_OK;
}

void impb_minphase_destroy(impb_minphase* mp) {
  if (!mp) return;
  DeviceGuard g(mp->ctx->device);
  mp->big.release();
  mp->small.release();
  cudaFree(mp->hamming);
  cudaFree(mp->nyq_S);
  cudaFree(mp->tab);
  cudaFree(mp->pos_a);
  cudaFree(mp->pos_b);
  cudaFree(mp->pos_j);
  delete mp;
}

static int64_t minphase_per_filter_bytes(const impb_minphase* mp) {
  // per filter (two filters share one complex transform): half a complex nf | two halves of complex numtaps |
  // real numtaps (h / cepstrum) | minpos | nyq imprint; rounded up for an odd filter count by the caller's + slack
  return mp->nf * 8 + mp->numtaps * 16 + mp->numtaps * 8 + 16 + (mp->nf / 2 + 1) * 8;  // + fx row
}

int64_t impb_minphase_workspace_bytes(const impb_minphase* mp, int64_t batch) {
  if (!mp || batch < 1) return 0;
  return (std::min<int64_t>(batch, mp->chunk) + 1) * minphase_per_filter_bytes(mp) + mp->numtaps * 8 + 256;
}

int32_t impb_minphase_fir(impb_minphase* mp, const double* gain, int64_t ld_gain, int32_t fs, double* fir_out,
                          int64_t ld_out, int64_t batch, void* workspace, void* stream) {
  if (!mp || !gain || !fir_out || !workspace || batch < 0 || ld_gain < mp->n || ld_out < mp->n || fs < 2) {
    set_error("impb_minphase_fir: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0) return IMPB_OK;
  DeviceGuard g(mp->ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t cb = std::min<int64_t>(batch, mp->chunk);  // filters per chunk (even, see impb_minphase_create)
  const int64_t cp = (cb + 1) / 2;                         // pairs per chunk
  char* w = (char*)workspace;
  double2* ws_big = (double2*)w;            w += cp * mp->nf * 16;
  double2* ws_a = (double2*)w;              w += cp * mp->numtaps * 16;
  double2* ws_b = (double2*)w;              w += cp * mp->numtaps * 16;
  double* real_buf = (double*)w;            w += 2 * cp * mp->numtaps * 8;
  double* minpos = (double*)w;              w += 2 * cp * 8;
  double* nyq_val = (double*)w;             w += 2 * cp * 8;
  double* fxn = (double*)w;                 w += 2 * cp * (mp->nf / 2 + 1) * 8;
  w = (char*)(((uintptr_t)w + 255) & ~(uintptr_t)255);
  double* hwin = (double*)w;
  static const bool fft_nyquist = getenv("IMPB_MINPHASE_FFT_NYQUIST") != nullptr;  // diagnostic: transform-computed bin
  const Plan2D& pb = mp->big.p;
  const Plan2D& ps = mp->small.p;
  const int64_t nt = mp->numtaps;
  const Mirror mir{mp->small.perm_col, mp->small.perm_row, mp->small.inv_col, mp->small.inv_row, ps.N1, ps.N2, ps.N};
  k_homomorphic_window<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(hwin, nt);
  for (int64_t b0 = 0; b0 < batch; b0 += cb) {
    const int nb = (int)std::min<int64_t>(cb, batch - b0);  // filters
    const int np = (nb + 1) / 2;                            // complex transforms
    int rc;
    // 1. firwin2: h = irfft(fx * shift)[:numtaps] * hamming, two filters per transform
    const double nyq = 0.5 * (double)fs;
    const int64_t nfreqs = mp->nf / 2 + 1;
    FirwinSpec spec{gain + (size_t)b0 * ld_gain, ld_gain, mp->n, mp->nf, nyq / (double)(nfreqs - 1), nyq, 1.0 / nyq,
                    (-(double)(nt - 1) / 2.0) * 3.141592653589793, (double)(fs / 2) / (double)(mp->n - 1),
                    (double)(fs / 2), (const FirwinTab*)mp->tab, mp->pos_a, mp->pos_b, mp->pos_j};
    if (mp->tab_fs != fs) {  // per-bin constants of this sampling rate (stream ordered; one rate per plan in practice)
      k_firwin_table<<<(unsigned)((nfreqs + 255) / 256), 256, 0, st>>>(spec, nt, mp->nyq_S, (FirwinTab*)mp->tab);
      k_firwin_pos_table<<<(unsigned)((mp->nf + 255) / 256), 256, 0, st>>>((const FirwinTab*)mp->tab, mp->nf, mp->big.perm_col,
                                                                        mp->big.perm_row, pb.N1, pb.N2, mp->pos_a, mp->pos_b,
                                                                        mp->pos_j);
      mp->tab_fs = fs;
    }
    // fx rows + imprint partial sums (in ws_a, free until step 2), then the imprints
    k_firwin_fx<<<dim3(IMPRINT_SLICES, nb), 256, 0, st>>>(spec, fxn, (double*)ws_a);
    k_mp_nyq_finish<<<(nb + 127) / 128, 128, 0, st>>>(spec, (const double*)ws_a, mp->win_sum, nb, nyq_val);
    static const bool inv_engine = getenv("IMPB_FIRWIN_INVERSE_ENGINE") != nullptr;  // A/B: the gathered DIT loader
    if (inv_engine) {
      RowInvFirwin2 ri{spec, nb, ws_big, mp->big.tw, pb.N, pb.N1, pb.N2, mp->big.perm_col, mp->big.perm_row};
      if ((rc = mp->big.row_inv<TILE>(ri, np, st))) return rc;
      ColInvReal2 c1{ws_big, pb.N, pb.N2, real_buf, nt, nt, nb, mp->hamming, 1.0 / (double)mp->nf};
      if ((rc = mp->big.col_inv<TILE>(c1, np, st))) return rc;
    } else {
      ColFwdFirwinFx2 cf{fxn, (const FirwinTab*)mp->tab, mp->nf, nb, ws_big, mp->big.tw, pb.N, pb.N2};
      if ((rc = mp->big.col_fwd<TILE>(cf, np, st))) return rc;
      RowFwdOp<RealOut2Op> rf1{ws_big, pb.N, pb.N1, pb.N2, mp->big.perm_col, mp->big.perm_row,
                               RealOut2Op{real_buf, nt, nt, nb, mp->hamming, 1.0 / (double)mp->nf}};
      if ((rc = mp->big.row_fwd<TILE>(rf1, np, st))) return rc;
    }
    // 2. Z = fft(h_a + i h_b, n_fft), kept in position order
    RealColIn2 ci{real_buf, nt, nt, nb, ws_a, mp->small.tw, ps.N, ps.N2};
    if ((rc = mp->small.col_fwd<TILE>(ci, np, st))) return rc;
    RowPlain<double, TILE> rf{ws_a, ws_a, nullptr, ps.N, ps.N1, ps.N2, 0, 0};
    if ((rc = mp->small.row_fwd<TILE>(rf, np, st))) return rc;
    const double* nv = fft_nyquist ? nullptr : nyq_val;
    k_fill_inf<<<(2 * np + 255) / 256, 256, 0, st>>>(minpos, 2 * np);
    k_min_positive_pair<<<dim3(8, np), 256, 0, st>>>(ws_a, mir, nv, nb, minpos);
    // 3. cepstra = ifft(0.5 log(|H| + 1e-7 min+)).real * homomorphic window
    RowInvPairLogMag rl{ws_a, minpos, nv, nb, mir, ws_b, mp->small.tw};
    if ((rc = mp->small.row_inv<TILE>(rl, np, st))) return rc;
    ColInvReal2 c2{ws_b, ps.N, ps.N2, real_buf, nt, nt, nb, hwin, 1.0 / (double)nt};
    if ((rc = mp->small.col_inv<TILE>(c2, np, st))) return rc;
    // 4. h_min = ifft(exp(fft(cepstrum))).real[:n]
    RealColIn2 ci2{real_buf, nt, nt, nb, ws_a, mp->small.tw, ps.N, ps.N2};
    if ((rc = mp->small.col_fwd<TILE>(ci2, np, st))) return rc;
    if ((rc = mp->small.row_fwd<TILE>(rf, np, st))) return rc;
    RowInvPairExp re{ws_a, nb, mir, ws_b, mp->small.tw};
    if ((rc = mp->small.row_inv<TILE>(re, np, st))) return rc;
    ColInvReal2 c3{ws_b, ps.N, ps.N2, fir_out + (size_t)b0 * ld_out, ld_out, mp->n, nb, nullptr, 1.0 / (double)nt};
    if ((rc = mp->small.col_inv<TILE>(c3, np, st))) return rc;
  }
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

int32_t impb_equalize(impb_ctx* ctx, const double* error, int64_t ld_err, const double* logf, const double* cap,
                      const double* gain_k, int64_t n, int32_t half_window, int32_t smoothen, double* out, int64_t ld_out,
                      int32_t* status, int64_t batch, void* stream) {
  if (!ctx || !error || !logf || !cap || !gain_k || !out || n < 1 || n > EQ_MAXF || ld_err < n || ld_out < n ||
      batch < 0 || half_window < 0) {
    set_error("impb_equalize: bad argument (n must be 1..%d)", EQ_MAXF);
    return IMPB_EINVAL;
  }
  if (batch == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  const size_t smem = EqShared::bytes((int)n);
  IMPB_CUDA(cudaFuncSetAttribute(k_equalize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EqShared::bytes(EQ_MAXF)));
  k_equalize<<<(unsigned)batch, 256, smem, (cudaStream_t)stream>>>(error, ld_err, logf, cap, gain_k, (int)n, half_window,
                                                                 smoothen, out, ld_out, status);
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

}  // extern "C"
