// K15 + K9 — decay window and polyphase resampling in one gather-MAC launch.
//   decay.apply_decay_window          core/decay.py:383-403
//   ImpulseResponse.resample          core/impulse_response.py:121-124
//     -> nnresample.resample -> scipy.signal.resample_poly(x, up, down, window=taps) -> upfirdn
// scipy's resample_poly (scipy/signal/_signaltools.py) pads the (already `up`-scaled) taps with
// n_pre_pad zeros, runs upfirdn and keeps y[n_pre_remove : n_pre_remove + n_out]; written out per output
// sample that is
//   out[j] = sum_k taps[(j + n_pre_remove) * down - n_pre_pad - k * up] * x[k]
// Each thread owns one output sample; the taps of one polyphase branch are read with stride `up`.
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

using namespace impb;

namespace {

struct DecayParams { double start, half, knee, level; };

__device__ __forceinline__ double decay_gain(const double* dp, int64_t k) {
  // window = concat(ones(start), hann(2*half)[half:], zeros(len - knee)) - 1;  10 ** (window * -level / 20)
  const double start = dp[0], half = dp[1], knee = dp[2], level = dp[3];
  double wv;
  if ((double)k < start) {
    wv = 1.0;
  } else if ((double)k < knee) {
    const double i = half + ((double)k - start);
    const double M = 2.0 * half;
    wv = 0.5 - 0.5 * cospi(2.0 * i / (M - 1.0));
  } else {
    wv = 0.0;
  }
  return exp10((wv - 1.0) * -level / 20.0);
}

__global__ void __launch_bounds__(256) k_resample(const double* __restrict__ src, int64_t ld_src, int64_t n,
                                                  const double* __restrict__ decay, const double* __restrict__ taps,
                                                  int64_t ntaps, int up, int down, int64_t n_pre_pad, int64_t n_pre_remove,
                                                  int64_t n_out, double* __restrict__ dst, int64_t ld_dst) {
  const int r = blockIdx.y;
  const double* x = src + (size_t)r * ld_src;
  const double* dp = decay ? decay + (size_t)r * 4 : nullptr;
  const bool use_decay = dp && !isnan(dp[0]);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_out; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = (j + n_pre_remove) * down - n_pre_pad;
    // tap index t - k*up in [0, ntaps)
    int64_t k_hi = t >= 0 ? t / up : -1;
    if (k_hi > n - 1) k_hi = n - 1;
    int64_t k_lo = t - (ntaps - 1);
    k_lo = k_lo > 0 ? (k_lo + up - 1) / up : 0;
    double acc = 0.0;
    // upfirdn accumulates with ascending tap index == descending k
    for (int64_t k = k_hi; k >= k_lo; --k) {
      double xv = x[k];
      if (use_decay) xv *= decay_gain(dp, k);
      acc += xv * taps[t - k * up];
    }
    dst[(size_t)r * ld_dst + j] = acc;
  }
}

// Register-tiled variant for up > 1: outputs j, j + up, j + 2 up ... share one polyphase branch (the same taps), and their
// input windows are `down` samples apart, so a thread that owns RT such outputs loads every tap once for RT multiply-adds
// (1 + RT loads per RT FMAs instead of 2 RT).  Each output is still one chain summed in ascending tap order, i.e. the
// values are those of k_resample.
template <int RT>
__global__ void __launch_bounds__(256) k_resample_tiled(const double* __restrict__ src, int64_t ld_src, int64_t n,
                                                        const double* __restrict__ decay, const double* __restrict__ taps,
                                                        int64_t ntaps, int up, int down, int64_t n_pre_pad,
                                                        int64_t n_pre_remove, int64_t n_out, double* __restrict__ dst,
                                                        int64_t ld_dst) {
  const int r = blockIdx.y;
  const double* x = src + (size_t)r * ld_src;
  const double* dp = decay ? decay + (size_t)r * 4 : nullptr;
  const bool use_decay = dp && !isnan(dp[0]);
  const int64_t groups = (n_out + (int64_t)RT * up - 1) / ((int64_t)RT * up);  // groups of RT * up consecutive outputs
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < groups * up; g += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j0 = (g / up) * ((int64_t)RT * up) + (g % up);
    if (j0 >= n_out) continue;
    const int64_t t0 = (j0 + n_pre_remove) * down - n_pre_pad;  // >= 0 for every output (n_pre_remove * down >= n_pre_pad)
    const int64_t phi = t0 % up, k0 = t0 / up;                   // taps phi + i * up pair with x[k0 + q * down - i]
    double acc[RT];
#pragma unroll
    for (int q = 0; q < RT; ++q) acc[q] = 0.0;
    const int64_t n_i = phi < ntaps ? (ntaps - 1 - phi) / up + 1 : 0;
    for (int64_t i = 0; i < n_i; ++i) {
      const double h = taps[phi + i * up];
#pragma unroll
      for (int q = 0; q < RT; ++q) {
        const int64_t k = k0 + (int64_t)q * down - i;
        if (k >= 0 && k < n) {
          double xv = x[k];
          if (use_decay) xv *= decay_gain(dp, k);
          acc[q] += xv * h;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < RT; ++q) {
      const int64_t j = j0 + (int64_t)q * up;
      if (j < n_out) dst[(size_t)r * ld_dst + j] = acc[q];
    }
  }
}

__global__ void __launch_bounds__(256) k_decay_only(const double* __restrict__ src, int64_t ld_src, int64_t n,
                                                    const double* __restrict__ decay, double* __restrict__ dst, int64_t ld_dst) {
  const int r = blockIdx.y;
  const double* dp = decay ? decay + (size_t)r * 4 : nullptr;
  const bool use_decay = dp && !isnan(dp[0]);
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    double v = src[(size_t)r * ld_src + k];
    if (use_decay) v *= decay_gain(dp, k);
    dst[(size_t)r * ld_dst + k] = v;
  }
}

}  // namespace

static int64_t gcd64(int64_t a, int64_t b) {
  while (b) { int64_t t = a % b; a = b; b = t; }
  return a;
}

extern "C" {

int64_t impb_resample_out_len(int64_t n, int32_t up, int32_t down) {
  if (n < 0 || up < 1 || down < 1) return -1;
  const int64_t g = gcd64(up, down);
  const int64_t u = up / g, d = down / g;
  return (n * u + d - 1) / d;
}

int32_t impb_window_resample(impb_ctx* ctx, const double* src, int64_t ld_src, int64_t n, const double* decay,
                             const double* taps, int64_t ntaps, int32_t up, int32_t down, double* dst, int64_t ld_dst,
                             int64_t batch, void* stream) {
  if (!ctx || !src || !dst || n < 0 || ld_src < n || batch < 0 || up < 1 || down < 1) {
    set_error("impb_window_resample: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0 || n == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t gg = gcd64(up, down);
  const int u = (int)(up / gg), d = (int)(down / gg);
  const int thr = 256;
  if (u == 1 && d == 1) {  // resample_poly returns a copy
    if (ld_dst < n) { set_error("impb_window_resample: ld_dst too small"); return IMPB_EINVAL; }
    const unsigned bx = (unsigned)std::min<int64_t>((n + thr - 1) / thr, 1024);
    for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
      const int nb = (int)std::min<int64_t>(65535, batch - b0);
      k_decay_only<<<dim3(bx, nb), thr, 0, st>>>(src + (size_t)b0 * ld_src, ld_src, n, decay ? decay + b0 * 4 : nullptr,
                                                dst + (size_t)b0 * ld_dst, ld_dst);
    }
    IMPB_CUDA(cudaGetLastError());
    return IMPB_OK;
  }
  if (!taps || ntaps < 1) { set_error("impb_window_resample: taps missing"); return IMPB_EINVAL; }
  const int64_t n_out = (n * u + d - 1) / d;
  if (ld_dst < n_out) { set_error("impb_window_resample: ld_dst < ceil(n*up/down)"); return IMPB_EINVAL; }
  const int64_t half_len = (ntaps - 1) / 2;
  const int64_t n_pre_pad = d - half_len % d;
  const int64_t n_pre_remove = (half_len + n_pre_pad) / d;
  const unsigned bx = (unsigned)std::min<int64_t>((n_out + thr - 1) / thr, 4096);
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = (int)std::min<int64_t>(65535, batch - b0);
    static const bool no_tile = getenv("IMPB_RESAMPLE_PLAIN") != nullptr;
    if (u > 1 && n_out >= 8 * (int64_t)u && !no_tile) {
      const unsigned bt = (unsigned)std::min<int64_t>(((n_out + 3) / 4 + thr - 1) / thr + 1, 4096);
      k_resample_tiled<4><<<dim3(bt, nb), thr, 0, st>>>(src + (size_t)b0 * ld_src, ld_src, n, decay ? decay + b0 * 4 : nullptr,
                                                        taps, ntaps, u, d, n_pre_pad, n_pre_remove, n_out,
                                                        dst + (size_t)b0 * ld_dst, ld_dst);
    } else {
      k_resample<<<dim3(bx, nb), thr, 0, st>>>(src + (size_t)b0 * ld_src, ld_src, n, decay ? decay + b0 * 4 : nullptr, taps,
                                              ntaps, u, d, n_pre_pad, n_pre_remove, n_out, dst + (size_t)b0 * ld_dst, ld_dst);
    }
  }
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

}  // extern "C"
