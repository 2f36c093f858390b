// Output packing: track-order gather -> PCM quantise -> interleave, on the device.
//   HRIR.write_wav          core/hrir.py:426-455   (rows in track order, missing tracks are silent)
//   audio_io.write_wav      core/audio_io.py:82-97 (transpose to frames x tracks, libsndfile PCM_16/24/32)
// Quantisation is libsndfile's double -> PCM conversion as the reference gets it through soundfile, which turns
// SFC_SET_CLIPPING on: every depth scales by 2^31, clips to the int32 range and takes lrint (round half to even)
// (pcm.c d2lei_clip_array); PCM_16 / PCM_24 then keep the TOP 16 / 24 bits of that value (d2les_clip_array /
// d2let_clip_array: value >> 16 / value >> 8, an arithmetic shift = floor), which is not a round at the target
// depth.  The reference's own committed golden data/demo/room-responses.wav (PCM_32) is reproduced bit for bit;
// no libsndfile-written PCM_16/24 vector exists in the reference tree (soundfile is not installed here), so the
// 16/24-bit rule is pinned to the library's published source only.
#include <algorithm>

#include "common.cuh"

using namespace impb;

namespace {

template <typename O>
__global__ void __launch_bounds__(256) k_pack_tracks(const double* __restrict__ src, const int64_t* __restrict__ row_offsets,
                                                     int64_t n, int n_tracks, int shift, O* __restrict__ dst) {
  // one thread per (frame, track); consecutive threads = consecutive tracks of a frame = consecutive output samples
  const int64_t total = n * n_tracks;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t frame = i / n_tracks;
    const int t = (int)(i - frame * n_tracks);
    const int64_t off = row_offsets[t];
    int32_t v = 0;
    if (off >= 0) {
      double q = rint(__dmul_rn(src[off + frame], 2147483648.0));
      q = fmin(fmax(q, -2147483648.0), 2147483647.0);
      v = (int32_t)q >> shift;
    }
    dst[i] = (O)v;
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
  const int shift = 32 - bit_depth;
  if (bit_depth == 16)
    k_pack_tracks<int16_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, row_offsets, n, n_tracks, shift, (int16_t*)dst);
  else
    k_pack_tracks<int32_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, row_offsets, n, n_tracks, shift, (int32_t*)dst);
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

}  // extern "C"
