// Device pieces of the virtual-bass synthesis (next row; reference core/virtual_bass.py:82-175).
//   impb_dft_bin_abs   _mag_at(): |rfft(x)[k]| of every row, evaluated as one direct DFT sum (any row length)
//   impb_add_delayed   _delay_signal() + the final  hi + synth  add: dst[r, j] += src_q[j - delay[r]]
// The IIR sections of the reference (scipy.signal.sosfilt) are applied as FFT convolutions with their own impulse
// responses by K1 (exact for the first n output samples of a causal filter); see virtual_bass.py.
#include <math.h>

#include <algorithm>

#include "common.cuh"

using namespace impb;

namespace {

constexpr int kDftThreads = 1024;

__global__ void __launch_bounds__(kDftThreads)
k_dft_bin_abs(const double* __restrict__ x, int64_t ld, int64_t n, int64_t k, double* __restrict__ out) {
  __shared__ double red[2 * (kDftThreads / 32)];
  const double* p = x + (size_t)blockIdx.x * ld;
  double re = 0.0, im = 0.0;
  // phase = 2 pi (k j mod n) / n, reduced in integers so that sincospi sees an argument in [0, 2).  The residue is
  // carried along the thread's stride (one 128-bit modulo per thread, then add-and-wrap).
  const uint64_t un = (uint64_t)n;
  uint64_t m = (uint64_t)(((unsigned __int128)(uint64_t)k * (uint64_t)threadIdx.x) % un);
  const uint64_t step = (uint64_t)(((unsigned __int128)(uint64_t)k * (uint64_t)kDftThreads) % un);
  const double inv_n = 2.0 / (double)n;
  for (int64_t j = threadIdx.x; j < n; j += kDftThreads) {
    double s, c;
    sincospi((double)m * inv_n, &s, &c);
    const double v = p[j];
    re += v * c;
    im -= v * s;
    m += step;
    if (m >= un) m -= un;
  }
  for (int o = 16; o; o >>= 1) {
    re += __shfl_xor_sync(0xffffffffu, re, o);
    im += __shfl_xor_sync(0xffffffffu, im, o);
  }
  if ((threadIdx.x & 31) == 0) { red[2 * (threadIdx.x >> 5)] = re; red[2 * (threadIdx.x >> 5) + 1] = im; }
  __syncthreads();
  if (threadIdx.x == 0) {
    re = im = 0.0;
    for (int w = 0; w < kDftThreads / 32; ++w) { re += red[2 * w]; im += red[2 * w + 1]; }
    out[blockIdx.x] = hypot(re, im);
  }
}

__global__ void __launch_bounds__(256)
k_add_delayed(double* __restrict__ dst, int64_t ld, int64_t n, const double* __restrict__ src, int64_t src_ld,
              int64_t src_len, const int32_t* __restrict__ which, const int64_t* __restrict__ delay) {
  const int64_t r = blockIdx.y;
  const double* s = src + (size_t)which[r] * src_ld;
  const int64_t d = delay[r];
  for (int64_t j = blockIdx.x * 256 + threadIdx.x; j < n; j += (int64_t)gridDim.x * 256) {
    const int64_t m = j - d;
    if (m >= 0 && m < src_len) dst[(size_t)r * ld + j] += s[m];
  }
}

}  // namespace

extern "C" {

int32_t impb_dft_bin_abs(impb_ctx* ctx, const double* x, int64_t ld, int64_t n, int64_t k, double* out, int64_t batch,
                         void* stream) {
  if (!ctx || !x || !out || batch < 0 || n <= 0 || ld < n || k < 0 || k >= n) {
    set_error("impb_dft_bin_abs: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int64_t nb = std::min<int64_t>(65535, batch - b0);
    k_dft_bin_abs<<<(unsigned)nb, kDftThreads, 0, st>>>(x + (size_t)b0 * ld, ld, n, k, out + b0);
  }
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

int32_t impb_add_delayed(impb_ctx* ctx, double* dst, int64_t ld, int64_t n, const double* src, int64_t src_ld,
                         int64_t src_len, int32_t n_src, const int32_t* which, const int64_t* delay, int64_t batch,
                         void* stream) {
  if (!ctx || !dst || !src || !which || !delay || batch < 0 || n < 0 || ld < n || src_len < 0 || src_ld < src_len ||
      n_src <= 0) {
    set_error("impb_add_delayed: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0 || n == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, 1024));
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int64_t nb = std::min<int64_t>(65535, batch - b0);
    k_add_delayed<<<dim3(gx, (unsigned)nb), 256, 0, st>>>(dst + (size_t)b0 * ld, ld, n, src, src_ld, src_len, which + b0,
                                                          delay + b0);
  }
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

}  // extern "C"
