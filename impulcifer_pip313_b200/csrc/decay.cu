// Decay analysis of cropped responses on the device (next row; decay_times also feeds
// decay_adjustment_params of row a11).
//   impb_decay_times     decay.decay_times (Schroeder backward integral, EDT / RT20 / RT30 / RT60)
//                        core/decay.py:263-352
//   impb_segment_energy  mean squares of index ranges, the arithmetic of HRIR.calculate_reflection_levels
//                        core/hrir.py:1003-1090
// One CTA per response.  The backward integral and the moving-average prefix sums are chunked scans: every
// thread owns a contiguous run of samples, run totals are scanned in shared memory, and the per-sample values
// go to a caller-provided scratch row so the regression passes can re-read them.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "exact_div.cuh"

using namespace impb;

namespace {

constexpr int DT_THREADS = 512;

struct DtShared {
  double red[DT_THREADS / 32];
  double run[DT_THREADS];
  double bcast[4];
  long long first[6];
};

__device__ __forceinline__ double block_max(double v, DtShared& sh) {
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh.red[threadIdx.x >> 5] = v;
  __syncthreads();
  double m = sh.red[0];
  for (int w = 1; w < DT_THREADS / 32; ++w) m = fmax(m, sh.red[w]);
  return m;
}

__device__ __forceinline__ double block_sum(double v, DtShared& sh) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh.red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < DT_THREADS / 32; ++w) s += sh.red[w];
  return s;
}

// slope of y[s:e) against t = i * dt, scipy.stats.linregress formulation (mean-centred sums, bias = 1)
__device__ double block_slope(const double* __restrict__ y, int64_t s, int64_t e, double dt, DtShared& sh) {
  const double n = (double)(e - s);
  double sx = 0.0, sy = 0.0;
  for (int64_t i = s + threadIdx.x; i < e; i += DT_THREADS) { sx += (double)i * dt; sy += y[i]; }
  const double xm = block_sum(sx, sh) / n;
  const double ym = block_sum(sy, sh) / n;
  double sxx = 0.0, sxy = 0.0;
  for (int64_t i = s + threadIdx.x; i < e; i += DT_THREADS) {
    const double dx = (double)i * dt - xm;
    sxx += dx * dx;
    sxy += dx * (y[i] - ym);
  }
  sxx = block_sum(sxx, sh);
  sxy = block_sum(sxy, sh);
  return (sxy / n) / (sxx / n);
}

__global__ void __launch_bounds__(DT_THREADS)
k_decay_times(const double* __restrict__ x, const int64_t* __restrict__ offsets, const int64_t* __restrict__ lengths,
              const int64_t* __restrict__ peak, const int64_t* __restrict__ knee_abs,
              const double* __restrict__ noise_db, const int64_t* __restrict__ win, int32_t fs,
              double* __restrict__ out, double* __restrict__ scratch, int64_t scratch_ld) {
  __shared__ DtShared sh;
  const int64_t r = blockIdx.x;
  const double* ir = x + offsets[r];
  const int64_t n = lengths[r], pk = peak[r], W = win[r];
  double* o = out + 4 * r;
  const int tid = threadIdx.x;
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
  if (tid < 4) o[tid] = nan;
  if (pk < 0 || pk >= n || W <= 0) return;
  const int64_t nd = n - pk;              // samples from the peak on
  const int64_t knee = knee_abs[r] - pk;  // knee relative to the peak
  const int64_t kk = min(knee, nd - 1);   // last sample of the backward integral; the curve has kk points
  if (kk <= 0) return;
  double* sdb = scratch + (size_t)r * scratch_ld;  // [kk] Schroeder curve in dB
  double* pre = sdb + (scratch_ld / 2);            // [La + 1] prefix sums of the moving average input
  const double* d = ir + pk;

  // normalised squares from the peak on, and the energy up to the knee
  double m = 0.0;
  for (int64_t j = tid; j < nd; j += DT_THREADS) m = fmax(m, fabs(d[j]));
  m = block_max(m, sh);
  const ExactDivisor by_m(m);  // bit-identical to d[j] / m (exact_div.cuh): one reciprocal per row
  double acc = 0.0;
  for (int64_t j = tid; j < min(knee, nd); j += DT_THREADS) { const double v = fabs(by_m(d[j])); acc += v * v; }
  const double denom = block_sum(acc, sh);
  const ExactDivisor by_denom(denom);

  // backward integral S[i] = sum_{j=i..kk} sq[j] / denom  (i < kk), chunked suffix scan
  const int64_t cnt = kk + 1;
  const int64_t C = (cnt + DT_THREADS - 1) / DT_THREADS;
  const int64_t lo = min((int64_t)tid * C, cnt), hi = min(lo + C, cnt);
  acc = 0.0;
  for (int64_t j = hi - 1; j >= lo; --j) { const double v = fabs(by_m(d[j])); acc += by_denom(v * v); }
  sh.run[tid] = acc;
  if (tid < 6) sh.first[tid] = (long long)kk;  // "not found"
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int t = DT_THREADS - 1; t >= 0; --t) { const double c = sh.run[t]; sh.run[t] = s; s += c; }
  }
  __syncthreads();
  {
    const double thr[6] = {-1.0, -5.0, -10.0, -25.0, -35.0, -65.0};
    long long fi[6] = {kk, kk, kk, kk, kk, kk};
    double s = sh.run[tid];
    for (int64_t j = hi - 1; j >= lo; --j) {
      const double v = fabs(by_m(d[j]));
      s += by_denom(v * v);
      if (j < kk) {
        const double db = 10.0 * log10(s);
        sdb[j] = db;
#pragma unroll
        for (int q = 0; q < 6; ++q)
          if (db <= thr[q]) fi[q] = j;  // descending j: the last assignment is the smallest index
      }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q)
      if (fi[q] < kk) atomicMin(&sh.first[q], fi[q]);
  }
  __syncthreads();

  // moving average of the squared, re-normalised response around [peak, knee] and the vertical offset that
  // lays the Schroeder curve over it
  const int64_t avg_head = min(W / 2, pk);
  const int64_t avg_tail = min(W / 2, n - (pk + knee));
  const int64_t avg_off = W / 2 - avg_head;
  const int64_t a0 = pk - avg_head, La = pk + knee + avg_tail - a0;
  double offset = nan;
  if (La >= W) {
    const double* a = ir + a0;
    double m2 = 0.0;
    for (int64_t j = tid; j < La; j += DT_THREADS) m2 = fmax(m2, fabs(a[j]));
    m2 = block_max(m2, sh);
    const ExactDivisor by_m2(m2);
    const int64_t C2 = (La + DT_THREADS - 1) / DT_THREADS;
    const int64_t l2 = min((int64_t)tid * C2, La), h2 = min(l2 + C2, La);
    acc = 0.0;
    for (int64_t j = l2; j < h2; ++j) { const double v = by_m2(a[j]); acc += v * v; }
    __syncthreads();
    sh.run[tid] = acc;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int t = 0; t < DT_THREADS; ++t) { const double c = sh.run[t]; sh.run[t] = s; s += c; }
    }
    __syncthreads();
    double s = sh.run[tid];
    if (tid == 0) pre[0] = 0.0;
    for (int64_t j = l2; j < h2; ++j) { const double v = by_m2(a[j]); s += v * v; pre[j + 1] = s; }
    __syncthreads();
    const int64_t len_avg = La - W + 1;
    const int64_t fit_s = max((int64_t)((double)kk * 0.1), avg_off);
    const int64_t fit_e = min((int64_t)((double)kk * 0.9), avg_off + len_avg);
    if (fit_e > fit_s) {
      acc = 0.0;
      for (int64_t i = fit_s + tid; i < fit_e; i += DT_THREADS) {
        const int64_t k = i - avg_off;
        const double rm = (pre[k + W] - pre[k]) / (double)W;
        acc += sdb[i] - 10.0 * log10(rm + 1e-18);
      }
      offset = block_sum(acc, sh) / (double)(fit_e - fit_s);
    }
  }
  __syncthreads();

  // EDT / RT20 / RT30 / RT60 from regression lines over the Schroeder curve
  const double dt = ((double)n / (double)fs) / (double)(n - 1);  // np.linspace(0, n / fs, n) spacing
  const int s_idx[4] = {0, 1, 1, 1}, e_idx[4] = {2, 3, 4, 5};
  const double e_thr[4] = {-10.0, -25.0, -35.0, -65.0}, target[4] = {-10.0, -20.0, -30.0, -60.0};
  const double nf = noise_db[r];
  for (int q = 0; q < 4; ++q) {
    if (e_thr[q] < nf + offset + 10.0) continue;  // less than 10 dB of headroom above the noise floor
    const long long s = sh.first[s_idx[q]], e = sh.first[e_idx[q]];
    if (s >= kk || e >= kk || s >= e || e - s < 2) continue;
    const double slope = block_slope(sdb, s, e, dt, sh);
    if (tid == 0) o[q] = target[q] / slope;
  }
}

// mean of squares of up to nseg index ranges per row: out[r * nseg + q] = mean(x[row r][lo:hi] ** 2), NaN if empty
__global__ void __launch_bounds__(256)
k_segment_energy(const double* __restrict__ x, const int64_t* __restrict__ offsets, const int64_t* __restrict__ seg_lo,
                 const int64_t* __restrict__ seg_hi, int nseg, double* __restrict__ out) {
  __shared__ double red[8];
  const int64_t r = blockIdx.x;
  const int q = blockIdx.y;
  const double* p = x + offsets[r];
  const int64_t lo = seg_lo[r * nseg + q], hi = seg_hi[r * nseg + q];
  double acc = 0.0;
  for (int64_t j = lo + threadIdx.x; j < hi; j += 256) acc += p[j] * p[j];
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w];
    out[r * nseg + q] = hi > lo ? s / (double)(hi - lo) : __longlong_as_double(0x7ff8000000000000ll);
  }
}

}  // namespace

extern "C" {

int64_t impb_decay_times_workspace_bytes(int64_t batch, int64_t max_len) {
  if (batch < 0 || max_len < 0) return 0;
  return batch * 2 * (max_len + 2) * (int64_t)sizeof(double);
}

int32_t impb_decay_times(impb_ctx* ctx, const double* x, const int64_t* offsets, const int64_t* lengths,
                         const int64_t* peak, const int64_t* knee, const double* noise_db, const int64_t* win, int32_t fs,
                         double* out, void* workspace, int64_t batch, int64_t max_len, void* stream) {
  if (!ctx || !x || !offsets || !lengths || !peak || !knee || !noise_db || !win || !out || !workspace || batch < 0 ||
      max_len < 0 || fs <= 0) {
    set_error("impb_decay_times: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int64_t nb = std::min<int64_t>(65535, batch - b0);
    k_decay_times<<<(unsigned)nb, DT_THREADS, 0, st>>>(x, offsets + b0, lengths + b0, peak + b0, knee + b0, noise_db + b0,
                                                       win + b0, fs, out + 4 * b0,
                                                       (double*)workspace + (size_t)b0 * 2 * (max_len + 2),
                                                       2 * (max_len + 2));
  }
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

int32_t impb_segment_energy(impb_ctx* ctx, const double* x, const int64_t* offsets, const int64_t* seg_lo,
                            const int64_t* seg_hi, int32_t nseg, double* out, int64_t batch, void* stream) {
  if (!ctx || !x || !offsets || !seg_lo || !seg_hi || !out || batch < 0 || nseg <= 0 || nseg > 65535) {
    set_error("impb_segment_energy: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int64_t nb = std::min<int64_t>(65535, batch - b0);
    k_segment_energy<<<dim3((unsigned)nb, (unsigned)nseg), 256, 0, st>>>(x, offsets + b0, seg_lo + b0 * nseg,
                                                                          seg_hi + b0 * nseg, nseg, out + b0 * nseg);
  }
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

}  // extern "C"
