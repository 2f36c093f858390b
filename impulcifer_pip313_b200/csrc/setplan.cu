// K22 — integer bookkeeping of a batch of recording sets between the device kernels (no host round trip).
//   HRIR.crop_heads                          core/hrir.py:548-612   pairwise first peak, speaker delay, crop index
//   HRIR.align_ipsilateral_all               core/hrir.py:921-955   lag -> which responses shift
//   HRIR.align_onset_groups_peak_leftref     core/hrir.py:957-1001  group onset difference -> shift of the group
//   HRIR.crop_tails                          core/hrir.py:614-653   tail = min(min len, next_fast_len(max knee))
// The reference does this arithmetic per recording set in Python; round 1 ran it as ~60 torch index / elementwise
// launches per batch plus one blocking device->host read in the middle of the step (the per-set tail had to reach
// the host before the final gather could be launched).  Here every stage is one small launch over all sets (one
// thread per response or per set), the per-set tails and the ragged output geometry are computed on the device, and
// the host reads the integers it reports only after the last launch of the step has been issued.
// All results are integers: bit-exact by construction (tests compare with the reference goldens).
#include <algorithm>

#include "common.cuh"

using namespace impb;

namespace {

__global__ void __launch_bounds__(256) k_set_heads(const impb_set_layout lay, const int64_t* __restrict__ peaks,
                                                   const int64_t* __restrict__ e_arg, int64_t L, int64_t seg, int64_t S,
                                                   int64_t* __restrict__ crop, int64_t* __restrict__ n1,
                                                   int64_t* __restrict__ off1, int64_t* __restrict__ hint_arg,
                                                   int64_t* __restrict__ xa_off, int64_t* __restrict__ xa_len,
                                                   int64_t* __restrict__ xb_off, int64_t* __restrict__ xb_len) {
  // phase 1: one thread per (set, speaker): crop index of both ears = max(0, min(peak_left, peak_right) - delay)
  const int64_t total_spk = S * lay.n_spk;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_spk; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / lay.n_spk;
    const int k = (int)(i % lay.n_spk);
    const int64_t rl = s * lay.n + lay.left[k], rr = s * lay.n + lay.right[k];
    const int64_t first = min(peaks[rl], peaks[rr]);
    const int64_t c = max(first - lay.delay[k], (int64_t)0);
    crop[rl] = c; crop[rr] = c;
    n1[rl] = L - c; n1[rr] = L - c;
    off1[rl] = rl * L + c; off1[rr] = rr * L + c;
    hint_arg[rl] = e_arg[rl] - c; hint_arg[rr] = e_arg[rr] - c;
  }
  if (!xa_off || lay.n_pairs == 0) return;
  // phase 2 needs phase 1 of the whole set: recompute the two crop indices it touches instead of a grid barrier
  const int64_t total_pairs = S * lay.n_pairs;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_pairs; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / lay.n_pairs;
    const int p = (int)(i % lay.n_pairs);
    const int64_t ra = s * lay.n + lay.pa[p], rb = s * lay.n + lay.pb[p];
    auto crop_of = [&](int row_in_set) -> int64_t {
      const int k = lay.spk_of_row[row_in_set];
      const int64_t first = min(peaks[s * lay.n + lay.left[k]], peaks[s * lay.n + lay.right[k]]);
      return max(first - lay.delay[k], (int64_t)0);
    };
    const int64_t ca = crop_of(lay.pa[p]), cb = crop_of(lay.pb[p]);
    xa_off[i] = ra * L + ca; xa_len[i] = min(L - ca, seg);
    xb_off[i] = rb * L + cb; xb_len[i] = min(L - cb, seg);
  }
}

// shift of every response from the ipsilateral lags, its surviving range, and the gathered views of the group leaders
__global__ void __launch_bounds__(256) k_set_shift(const impb_set_layout lay, const int64_t* __restrict__ lags,
                                                   const int64_t* __restrict__ n1, const int64_t* __restrict__ off1,
                                                   const int64_t* __restrict__ hint_arg, const double* __restrict__ e_max,
                                                   int64_t S, int64_t* __restrict__ shift, int64_t* __restrict__ lo,
                                                   int64_t* __restrict__ hi, int64_t* __restrict__ g_off,
                                                   int64_t* __restrict__ g_len, int64_t* __restrict__ g_shift,
                                                   int64_t* __restrict__ g_lo, int64_t* __restrict__ g_hi,
                                                   double* __restrict__ g_max, int64_t* __restrict__ g_arg) {
  const int64_t s = blockIdx.x;
  if (s >= S) return;
  extern __shared__ long long sh_shift[];  // [n]
  for (int r = threadIdx.x; r < lay.n; r += blockDim.x) sh_shift[r] = 0;
  __syncthreads();
  if (lags && threadIdx.x == 0) {  // a handful of entries per set, in the reference's order (integer adds commute)
    const int64_t* lg = lags + s * lay.n_pairs;
    for (int i = 0; i < lay.n_pos; ++i) sh_shift[lay.pos_rows[i]] += max(lg[lay.pos_k[i]], (int64_t)0);
    for (int i = 0; i < lay.n_neg; ++i) sh_shift[lay.neg_rows[i]] += max(-lg[lay.neg_k[i]], (int64_t)0);
  }
  __syncthreads();
  for (int r = threadIdx.x; r < lay.n; r += blockDim.x) {
    const int64_t row = s * lay.n + r, sv = sh_shift[r], len = n1[row];
    const int64_t l = max(sv, (int64_t)0);
    shift[row] = sv;
    lo[row] = l;
    hi[row] = max(min(len + sv, len), l);
  }
  __syncthreads();
  if (!g_off) return;
  for (int g = threadIdx.x; g < lay.n_groups; g += blockDim.x) {
    const int64_t row = s * lay.n + lay.gl[g], o = s * lay.n_groups + g;
    const int64_t sv = sh_shift[lay.gl[g]], len = n1[row], l = max(sv, (int64_t)0);
    g_off[o] = off1[row]; g_len[o] = len; g_shift[o] = sv; g_lo[o] = l; g_hi[o] = max(min(len + sv, len), l);
    g_max[o] = e_max[row]; g_arg[o] = hint_arg[row];
  }
}

// group onset alignment: every response of a group moves by -(leader peak - FL-left peak); composite view
__global__ void __launch_bounds__(256) k_set_shift2(const impb_set_layout lay, const int64_t* __restrict__ gp,
                                                    const int64_t* __restrict__ n1, int64_t S, int64_t* __restrict__ shift,
                                                    int64_t* __restrict__ lo, int64_t* __restrict__ hi) {
  const int64_t s = blockIdx.x;
  if (s >= S) return;
  extern __shared__ long long sh2[];  // [n]
  for (int r = threadIdx.x; r < lay.n; r += blockDim.x) sh2[r] = 0;
  __syncthreads();
  if (gp && threadIdx.x == 0) {
    const int64_t* g = gp + s * lay.n_groups;
    const int64_t ref = g[lay.g_ref];
    for (int i = 0; i < lay.n_grows; ++i) sh2[lay.g_rows[i]] += -(g[lay.g_of_row[i]] - ref);
  }
  __syncthreads();
  for (int r = threadIdx.x; r < lay.n; r += blockDim.x) {
    const int64_t row = s * lay.n + r, s2 = sh2[r], len = n1[row];
    const int64_t l = max(lo[row] + s2, (int64_t)0);
    shift[row] += s2;  // two length-preserving shifts compose into one shift plus a surviving range
    lo[row] = l;
    hi[row] = max(min(hi[row] + s2, len), l);
  }
}

// per-set tail and the ragged output geometry.  One CTA; sets are walked by threads, offsets by a block scan.
__global__ void __launch_bounds__(1024) k_set_tails(const impb_set_layout lay, const int64_t* __restrict__ knee,
                                                    const int64_t* __restrict__ n1, int64_t S,
                                                    const int64_t* __restrict__ smooth, int n_smooth, int64_t head,
                                                    int64_t fade_out_len, int64_t tail_bound, int64_t* __restrict__ tails,
                                                    int64_t* __restrict__ set_off, int64_t* __restrict__ out_len,
                                                    int64_t* __restrict__ dst_off, int64_t* __restrict__ flags) {
  __shared__ long long part[1024];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  long long f_short = 0, f_fade = 0, f_bound = 0;
  for (int64_t base = 0; base < S; base += blockDim.x) {
    const int64_t s = base + threadIdx.x;
    long long mine = 0;
    int64_t tail = 0;
    if (s < S) {
      int64_t mn = INT64_MAX, mk = 0;
      for (int r = 0; r < lay.n; ++r) {
        mn = min(mn, n1[s * lay.n + r]);
        mk = max(mk, knee[s * lay.n + r]);
      }
      // scipy.fftpack.next_fast_len(target): smallest 5-smooth number >= target (table built by the host shim)
      int a = 0, b = n_smooth - 1;
      while (a < b) {
        const int m = (a + b) >> 1;
        if (smooth[m] >= mk) b = m; else a = m + 1;
      }
      int64_t nfl = smooth[a];
      if (nfl < mk) { nfl = mk; f_bound = 1; }  // beyond the table: report, never guess
      tail = min(mn, nfl);
      if (head > 0 && mn < head) f_short = 1;
      if (tail < fade_out_len) f_fade = 1;
      if (tail > tail_bound) { f_bound = 1; tail = tail_bound; }
      tails[s] = tail;
      mine = (long long)tail * lay.n;
    }
    part[threadIdx.x] = mine;
    __syncthreads();
    for (int o = 1; o < (int)blockDim.x; o <<= 1) {  // inclusive Hillis-Steele scan
      long long v = threadIdx.x >= (unsigned)o ? part[threadIdx.x - o] : 0;
      __syncthreads();
      part[threadIdx.x] += v;
      __syncthreads();
    }
    const long long excl = carry + part[threadIdx.x] - mine;
    if (s < S) {
      set_off[s] = excl;
      for (int r = 0; r < lay.n; ++r) {
        out_len[s * lay.n + r] = tail;
        dst_off[s * lay.n + r] = excl + (long long)r * tail;
      }
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += part[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) set_off[S] = carry;
  if (f_short) atomicOr((unsigned long long*)flags, 1ull);
  if (f_fade) atomicOr((unsigned long long*)flags, 2ull);
  if (f_bound) atomicOr((unsigned long long*)flags, 4ull);
}

}  // namespace

extern "C" {

int32_t impb_set_heads(impb_ctx* ctx, const impb_set_layout* lay, const int64_t* peaks, const int64_t* e_arg, int64_t L,
                       int64_t seg, int64_t S, int64_t* crop, int64_t* n1, int64_t* off1, int64_t* hint_arg,
                       int64_t* xa_off, int64_t* xa_len, int64_t* xb_off, int64_t* xb_len, void* stream) {
  if (!ctx || !lay || !peaks || !e_arg || !crop || !n1 || !off1 || !hint_arg || S < 0 || lay->n < 1 || lay->n_spk < 1) {
    set_error("impb_set_heads: bad argument");
    return IMPB_EINVAL;
  }
  if (S == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  const int64_t work = S * std::max(lay->n_spk, lay->n_pairs);
  k_set_heads<<<(unsigned)std::min<int64_t>((work + 255) / 256, 4096), 256, 0, (cudaStream_t)stream>>>(
      *lay, peaks, e_arg, L, seg, S, crop, n1, off1, hint_arg, xa_off, xa_len, xb_off, xb_len);
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

int32_t impb_set_shift(impb_ctx* ctx, const impb_set_layout* lay, const int64_t* lags, const int64_t* n1,
                       const int64_t* off1, const int64_t* hint_arg, const double* e_max, int64_t S, int64_t* shift,
                       int64_t* lo, int64_t* hi, int64_t* g_off, int64_t* g_len, int64_t* g_shift, int64_t* g_lo,
                       int64_t* g_hi, double* g_max, int64_t* g_arg, void* stream) {
  if (!ctx || !lay || !n1 || !shift || !lo || !hi || S < 0 || (g_off && (!off1 || !hint_arg || !e_max))) {
    set_error("impb_set_shift: bad argument");
    return IMPB_EINVAL;
  }
  if (S == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  k_set_shift<<<(unsigned)S, 64, sizeof(long long) * lay->n, (cudaStream_t)stream>>>(
      *lay, lags, n1, off1, hint_arg, e_max, S, shift, lo, hi, g_off, g_len, g_shift, g_lo, g_hi, g_max, g_arg);
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

int32_t impb_set_shift2(impb_ctx* ctx, const impb_set_layout* lay, const int64_t* group_peaks, const int64_t* n1, int64_t S,
                        int64_t* shift, int64_t* lo, int64_t* hi, void* stream) {
  if (!ctx || !lay || !n1 || !shift || !lo || !hi || S < 0 || (group_peaks && (lay->g_ref < 0 || lay->g_ref >= lay->n_groups))) {
    set_error("impb_set_shift2: bad argument");
    return IMPB_EINVAL;
  }
  if (S == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  k_set_shift2<<<(unsigned)S, 64, sizeof(long long) * lay->n, (cudaStream_t)stream>>>(*lay, group_peaks, n1, S, shift, lo, hi);
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

int32_t impb_set_tails(impb_ctx* ctx, const impb_set_layout* lay, const int64_t* knee, const int64_t* n1, int64_t S,
                       const int64_t* smooth, int32_t n_smooth, int64_t head, int64_t fade_out_len, int64_t tail_bound,
                       int64_t* tails, int64_t* set_off, int64_t* out_len, int64_t* dst_off, int64_t* flags, void* stream) {
  if (!ctx || !lay || !knee || !n1 || !smooth || n_smooth < 1 || !tails || !set_off || !out_len || !dst_off || !flags ||
      S < 0) {
    set_error("impb_set_tails: bad argument");
    return IMPB_EINVAL;
  }
  DeviceGuard g(ctx->device);
  IMPB_CUDA(cudaMemsetAsync(flags, 0, sizeof(int64_t), (cudaStream_t)stream));
  k_set_tails<<<1, 1024, 0, (cudaStream_t)stream>>>(*lay, knee, n1, S, smooth, n_smooth, head, fade_out_len, tail_bound,
                                                    tails, set_off, out_len, dst_off, flags);
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

}  // extern "C"
