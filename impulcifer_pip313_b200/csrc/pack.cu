// Output packing: track-order gather -> PCM quantise -> interleave, on the device.
//   HRIR.write_wav          core/hrir.py:426-455   (rows in track order, missing tracks are silent)
//   audio_io.write_wav      core/audio_io.py:82-97 (transpose to frames x tracks, libsndfile PCM_16/24/32)
// Quantisation is libsndfile's float -> PCM conversion as the reference gets it through soundfile:
// lrint(x * 2^(bits-1)) (round half to even) with clipping to the integer range; the reference's own committed
// golden data/demo/room-responses.wav is reproduced bit for bit with this rule.
#include <algorithm>

#include "common.cuh"

using namespace impb;

namespace {

template <typename O>
__global__ void __launch_bounds__(256) k_pack_tracks(const double* __restrict__ src, const int64_t* __restrict__ row_offsets,
                                                     int64_t n, int n_tracks, double full, O* __restrict__ dst) {
  // one thread per (frame, track); consecutive threads = consecutive tracks of a frame = consecutive output samples
  const int64_t total = n * n_tracks;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t frame = i / n_tracks;
    const int t = (int)(i - frame * n_tracks);
    const int64_t off = row_offsets[t];
    double q = 0.0;
    if (off >= 0) {
      q = rint(__dmul_rn(src[off + frame], full));
      q = fmin(fmax(q, -full), full - 1.0);
    }
    dst[i] = (O)q;
  }
}

}  // namespace

extern "C" {

int32_t impb_pack_tracks(impb_ctx* ctx, const double* src, const int64_t* row_offsets, int64_t n, int32_t n_tracks,
                         int32_t bit_depth, void* dst, void* stream) {
  if (!ctx || !src || !row_offsets || !dst || n < 0 || n_tracks < 1 ||
      (bit_depth != 16 && bit_depth != 24 && bit_depth != 32)) {
    set_error(bit_depth != 16 && bit_depth != 24 && bit_depth != 32
                  ? "Invalid bit depth. Accepted values are 16, 24 and 32."
                  : "impb_pack_tracks: bad argument");
    return IMPB_EINVAL;
  }
  if (n == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  const int64_t total = n * n_tracks;
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 65536);
  const double full = bit_depth == 16 ? 32768.0 : bit_depth == 24 ? 8388608.0 : 2147483648.0;
  if (bit_depth == 16)
    k_pack_tracks<int16_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, row_offsets, n, n_tracks, full, (int16_t*)dst);
  else
    k_pack_tracks<int32_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, row_offsets, n, n_tracks, full, (int32_t*)dst);
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

}  // extern "C"
