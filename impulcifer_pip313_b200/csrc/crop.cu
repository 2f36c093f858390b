// K2 / K4 / K6 / K3+K5+K7+K16: integer logic of the crop / align path on the device.
//   impb_peak_index    ImpulseResponse.peak_index      core/impulse_response.py:32-70 (twin core/decay.py:12-41)
//   impb_lundeby       decay.decay_params              core/decay.py:44-260
//   impb_xcorr_lag     correlate + argmax              core/hrir.py:932-948
//   impb_gather_window crop / shift / fade / gain      core/hrir.py:591-612, 642-651, 527-544;
//                                                      core/impulse_response.py:92-108
// All index results are exact integers: comparisons are done on the same IEEE operations the reference
// performs (x / max|x| by true division, no FMA contraction on the values that are compared).
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "exact_div.cuh"

using namespace impb;

namespace {

constexpr double kEps = 1e-20;
constexpr unsigned long long kNoIndex = 0xFFFFFFFFFFFFFFFFull;

// A "view" of one response: the array the reference would hold after head crop, fade-in and a chain of
// length-preserving shifts, expressed as index arithmetic over the untouched source (no materialised copy):
//   v[j] = fade(j - shift) * src[start + j - shift]   for lo <= j < hi,   0 elsewhere      (0 <= j < n)
struct ViewArgs {
  const int64_t* offsets;  // start of row r in x (after the head crop)
  const int64_t* lengths;  // logical length n of row r
  const int64_t* shift;    // nullable
  const int64_t* lo;       // nullable
  const int64_t* hi;       // nullable
  const double* fade;      // nullable: rising half window applied to source index m < fade_len
  int64_t fade_len;
  const double* hint_max;      // nullable pair: max |src| over a superset of the view and one source index
  const int64_t* hint_arg;     //                (relative to offsets[r]) where it is attained
  const double* blk_max;       // nullable: [rows][nblk] block maxima of the untouched samples
  int64_t nblk, blk_len, blk_first;
};

struct RowView {
  const double* p;
  int64_t n, sh, lo, hi, fade_len;
  const double* fade;
  __device__ __forceinline__ RowView(const double* x, const ViewArgs& a, int64_t r) {
    p = x + a.offsets[r];
    n = a.lengths[r];
    sh = a.shift ? a.shift[r] : 0;
    lo = a.lo ? a.lo[r] : 0;
    hi = a.hi ? a.hi[r] : n;
    fade = a.fade;
    fade_len = a.fade ? a.fade_len : 0;
  }
  // view index of the hinted maximum if that sample is part of the view with weight 1, else -1
  __device__ __forceinline__ int64_t hinted(const ViewArgs& a, int64_t r) const {
    if (!a.hint_max || !a.hint_arg) return -1;
    const int64_t m = a.hint_arg[r], j = m + sh;
    return (m >= fade_len && j >= lo && j < hi && j >= 0 && j < n) ? j : -1;
  }
  __device__ __forceinline__ double operator[](int64_t j) const {
    if (j < lo || j >= hi) return 0.0;
    const int64_t m = j - sh;
    double v = p[m];
    if (m < fade_len) v = __dmul_rn(v, fade[m]);
    return v;
  }
};

struct PeakScratch {
  unsigned long long max_bits;  // bits of max |x| (non-negative doubles order like integers)
  unsigned long long first_peak;
  unsigned long long first_argmax;
};

__global__ void k_peak_init(PeakScratch* s, int64_t batch) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < batch) {
    s[i].max_bits = 0ull;
    s[i].first_peak = kNoIndex;
    s[i].first_argmax = kNoIndex;
  }
}

// grid = (rows, chunks): block (r, k) owns the contiguous samples [k*C, (k+1)*C) of row r.  Chunk index is the
// SLOW grid dimension, so blocks are scheduled front-to-back over all rows and the peak search below can drop
// every chunk behind an already published peak with a single read.
__device__ __forceinline__ void chunk_bounds(int64_t n, int64_t* lo, int64_t* hi) {
  int64_t C = (n + gridDim.y - 1) / gridDim.y;
  C = (C + 255) & ~(int64_t)255;
  *lo = (int64_t)blockIdx.y * C;
  *hi = min(n, *lo + C);
}

__global__ void __launch_bounds__(256) k_absmax(const double* __restrict__ x, const ViewArgs va, int64_t r0,
                                                PeakScratch* __restrict__ s) {
  const int64_t r = blockIdx.x;
  const RowView p(x, va, r0 + r);
  if (p.hinted(va, r0 + r) >= 0) {  // the known maximum of the source survives in this view: no scan
    if (blockIdx.y == 0 && threadIdx.x == 0)
      s[r].max_bits = (unsigned long long)__double_as_longlong(va.hint_max[r0 + r]);
    return;
  }
  // grid = (rows, G): block g streams the 256-aligned slices g, g + G, ... of the row (few, fat blocks: the rows that
  // take their maximum from the hint cost one block launch each instead of one per 4096 samples)
  double m = 0.0;
  const int64_t slice = 4096;
  for (int64_t lo = (int64_t)blockIdx.y * slice; lo < p.n; lo += (int64_t)gridDim.y * slice) {
    const int64_t hi = min(p.n, lo + slice);
    int64_t i = lo + threadIdx.x;
    for (; i + 768 < hi; i += 1024) {  // four independent loads in flight per thread
      const double a0 = p[i], a1 = p[i + 256], a2 = p[i + 512], a3 = p[i + 768];
      m = fmax(fmax(m, fmax(fabs(a0), fabs(a1))), fmax(fabs(a2), fabs(a3)));
    }
    for (; i < hi; i += 256) m = fmax(m, fabs(p[i]));
  }
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ double sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmax(m, sm[w]);
    atomicMax(&s[r].max_bits, (unsigned long long)__double_as_longlong(m));
  }
}

// candidate test for sample i of the normalised signal q = sign * x / m (scipy _local_maxima_1d semantics)
__device__ __forceinline__ bool peak_at(const RowView& p, int64_t n, int64_t i, double m, double sign, double height,
                                        int64_t* mid) {
  const double qi = sign * __ddiv_rn(p[i], m);
  if (!(qi >= height)) return false;
  const double qp = sign * __ddiv_rn(p[i - 1], m);
  if (!(qp < qi)) return false;
  int64_t j = i + 1;
  const int64_t i_max = n - 1;
  double qj = sign * __ddiv_rn(p[j], m);
  while (j < i_max && qj == qi) {
    ++j;
    qj = sign * __ddiv_rn(p[j], m);
  }
  if (!(qj < qi)) return false;
  *mid = (i + (j - 1)) / 2;
  return true;
}

// grid = (rows, G): block (r, g) walks the chunks g, g + G, g + 2G ... of row r front to back and stops as soon as an
// earlier chunk has published a peak, so a row whose onset lies in the first G chunks costs G blocks, not `nchunks`.
__global__ void __launch_bounds__(256) k_first_peak(const double* __restrict__ x, const ViewArgs va, int64_t r0,
                                                    double height, PeakScratch* __restrict__ s, int nchunks) {
  const int64_t r = blockIdx.x;
  const RowView p(x, va, r0 + r);
  const int64_t n = p.n;
  const double m = __longlong_as_double((long long)s[r].max_bits);
  if (!(m >= kEps)) return;
  int64_t C = (n + nchunks - 1) / nchunks;
  C = (C + 255) & ~(int64_t)255;
  const double guard = 0.999 * height * m;
  // a peak candidate is a sample >= guard: nothing before the first block whose maximum reaches it can qualify
  int64_t k0 = 0;
  if (va.blk_max && guard <= m) {  // guard > max|x| (peak_height > ~1.001): no candidate exists, argmax|x| is the
                                   // answer and it may sit anywhere, so no block may be skipped
    const double* bm = va.blk_max + (r0 + r) * va.nblk;
    int64_t skip = n;
    for (int64_t q = 0; q < va.nblk; ++q)
      if (bm[q] >= guard) { skip = va.blk_first + q * va.blk_len; break; }
    k0 = max((int64_t)0, min(skip, n - 1)) / C;
  }
  for (int64_t k = k0 + blockIdx.y; k < nchunks; k += gridDim.y) {
    const int64_t lo = k * C, hi = min(n, lo + C);
    if (lo >= hi) return;
    // nothing in this chunk (or any later one) can beat an already published earlier peak
    // (first_argmax is only consulted when no peak exists at all, so it need not be complete then)
    const unsigned long long cur = *((volatile unsigned long long*)&s[r].first_peak);
    if (cur != kNoIndex && cur < (unsigned long long)lo) return;
    auto visit = [&](int64_t i, double v) {
      const double av = fabs(v);
      if (av == m) atomicMin(&s[r].first_argmax, (unsigned long long)i);
      if (i < 1 || i > n - 2 || av < guard) return;
      int64_t mid;
      if (peak_at(p, n, i, m, v > 0 ? 1.0 : -1.0, height, &mid)) atomicMin(&s[r].first_peak, (unsigned long long)mid);
    };
    int64_t i = lo + threadIdx.x;
    if (lo >= max(p.lo, p.fade_len + p.sh) && hi <= p.hi) {
      // the chunk lies in the unfaded valid part of the view: eight plain loads per thread, nothing between them
      const double* q = p.p - p.sh;
      for (; i + 1792 < hi; i += 2048) {
        double a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = q[i + u * 256];
#pragma unroll
        for (int u = 0; u < 8; ++u) visit(i + u * 256, a[u]);
      }
    }
    for (; i + 768 < hi; i += 1024) {  // four independent loads in flight per thread
      const double a0 = p[i], a1 = p[i + 256], a2 = p[i + 512], a3 = p[i + 768];
      visit(i, a0); visit(i + 256, a1); visit(i + 512, a2); visit(i + 768, a3);
    }
    for (; i < hi; i += 256) visit(i, p[i]);
  }
}

__global__ void k_peak_final(const PeakScratch* __restrict__ s, int64_t* __restrict__ out, int64_t batch) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= batch) return;
  const double m = __longlong_as_double((long long)s[i].max_bits);
  int64_t r = 0;
  if (m >= kEps) r = s[i].first_peak != kNoIndex ? (int64_t)s[i].first_peak : (int64_t)s[i].first_argmax;
  out[i] = r;
}

__global__ void k_peak_seed(PeakScratch* s, const double* __restrict__ known_max, int64_t batch) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < batch) {
    s[i].max_bits = (unsigned long long)__double_as_longlong(known_max[i]);
    s[i].first_peak = kNoIndex;
    s[i].first_argmax = kNoIndex;
  }
}

int run_peak_index(impb_ctx* ctx, const double* x, const ViewArgs& va, double height, int64_t* peak_out, int64_t batch,
                   int64_t max_len_hint, cudaStream_t st, const double* known_max = nullptr) {
  PeakScratch* scratch = nullptr;
  IMPB_CUDA(cudaMallocAsync(&scratch, sizeof(PeakScratch) * batch, st));
  const int thr = 256;
  if (known_max)
    k_peak_seed<<<(unsigned)((batch + thr - 1) / thr), thr, 0, st>>>(scratch, known_max, batch);
  else
    k_peak_init<<<(unsigned)((batch + thr - 1) / thr), thr, 0, st>>>(scratch, batch);
  // enough blocks per segment to fill the machine for small batches, few for large ones
  // ~16 samples per thread
  static const int spt = getenv("IMPB_PEAK_SPT") ? atoi(getenv("IMPB_PEAK_SPT")) : 16;
  int64_t chunks = std::max<int64_t>(1, (max_len_hint + thr * spt - 1) / (thr * spt));
  chunks = std::min<int64_t>(chunks, 4096);
  for (int64_t b0 = 0; b0 < batch; b0 += (int64_t)1 << 30) {
    const int64_t nb = std::min<int64_t>((int64_t)1 << 30, batch - b0);
    dim3 grid((unsigned)nb, (unsigned)chunks);
    if (!known_max) {
      // with hints most rows return at once: few blocks per row; without, one block per slice streams best
      const int64_t mg = va.hint_max ? std::max<int64_t>(8, 4096 / std::max<int64_t>(nb, 1)) : chunks;
      const dim3 mgrid((unsigned)nb, (unsigned)std::min<int64_t>(chunks, mg));
      k_absmax<<<mgrid, thr, 0, st>>>(x, va, b0, scratch + b0);
    }
    // Blocks per row.  Head-cropped views (they carry the max hint) have their onset within the first samples: the
    // row is walked front to back by just enough blocks to fill the machine (896 rows on B200: 2 blocks per row
    // 111 us, one block per chunk 192 us, mostly block scheduling).  Raw estimates have the onset anywhere (the
    // 'same' slice puts it ~40 % into the row): every chunk gets its own block and the ones behind a published
    // peak drop out (2 blocks per row: 1004 us, one per chunk: ~200 us).
    static const int peak_g = getenv("IMPB_PEAK_G") ? atoi(getenv("IMPB_PEAK_G")) : 0;
    const bool near_start = va.hint_max != nullptr || va.blk_max != nullptr;
    const int64_t want = peak_g > 0 ? peak_g : near_start ? std::max<int64_t>(1, 2048 / std::max<int64_t>(nb, 1)) : chunks;
    const dim3 pgrid((unsigned)nb, (unsigned)std::min<int64_t>(chunks, want));
    k_first_peak<<<pgrid, thr, 0, st>>>(x, va, b0, height, scratch + b0, (int)chunks);
  }
  k_peak_final<<<(unsigned)((batch + thr - 1) / thr), thr, 0, st>>>(scratch, peak_out, batch);
  IMPB_CUDA(cudaGetLastError());
  IMPB_CUDA(cudaFreeAsync(scratch, st));
  return IMPB_OK;
}

// ------------------------------------------------------------------------------------------------
// Lundeby: one CTA per impulse response
// ------------------------------------------------------------------------------------------------
#ifndef IMPB_LB_THREADS
#define IMPB_LB_THREADS 256  // two CTAs per SM at 512 threads (registers): 256 keeps three responses per SM in flight (-6 %)
#endif
constexpr int LB_THREADS = IMPB_LB_THREADS;
constexpr int LB_MAXWIN = 4096;  // >= the 4000-window upper bound of the refined pass (proof at its use)

constexpr int LB_MAXBLK = 3072;  // partial sums of the squared window: blocks of bs = 32, 64, ... samples

struct LundebyShared {
  double red[LB_THREADS / 32];
  double bcast[4];
  long long ibcast[4];
  double win[LB_MAXWIN];
  double blk[LB_MAXBLK];
};

__device__ __forceinline__ double block_sum(double v, LundebyShared& sh) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh.red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int w = 0; w < LB_THREADS / 32; ++w) t += sh.red[w];
  return t;
}

struct SqView {  // squared, peak-normalised analysis segment  sq[i] = (v[pk+i] / mx)^2
  const RowView* p;
  int64_t pk;
  double mx;
  bool norm;
  const double* blk;  // shared-memory sums of aligned blocks of `bs` samples (nullptr: not built)
  int bs;
  ExactDivisor div;   // division by mx with the reciprocal rounded once per row
  __device__ __forceinline__ double of(double s) const {
    const double v = norm ? div(s) : s;  // == __ddiv_rn(s, mx), see exact_div.cuh
    return __dmul_rn(v, v);
  }
  __device__ __forceinline__ double operator()(int64_t i) const { return of((*p)[pk + i]); }
};

// Every statistic of decay_params is a mean of sq over a contiguous range.  The samples are read (and divided)
// ONCE into per-block partial sums (k_lb_blocksums, copied to shared memory by k_lundeby); a range sum is then whole
// blocks from shared memory plus at most 2*bs edge samples from HBM.  No prefix-sum differences are taken, so there
// is no cancellation.
// partial (per-thread) sum of sq over [a, b) for threads `tid` of `nthr` cooperating threads
__device__ __forceinline__ double range_partial(const SqView& sq, int64_t a, int64_t b, int tid, int nthr) {
  double acc = 0.0;
  if (!sq.blk) {
    for (int64_t i = a + tid; i < b; i += nthr) acc += sq(i);
    return acc;
  }
  const int64_t bs = sq.bs;
  const int64_t a_up = min(b, (a + bs - 1) / bs * bs);
  const int64_t b_dn = max(a_up, b / bs * bs);
  for (int64_t i = a + tid; i < a_up; i += nthr) acc += sq(i);
  for (int64_t k = a_up / bs + tid; k < b_dn / bs; k += nthr) acc += sq.blk[k];
  for (int64_t i = b_dn + tid; i < b; i += nthr) acc += sq(i);
  return acc;
}

__device__ __forceinline__ double range_mean(const SqView& sq, int64_t a, int64_t b, LundebyShared& sh) {
  return block_sum(range_partial(sq, a, b, threadIdx.x, LB_THREADS), sh) / (double)(b - a);
}

__device__ __forceinline__ double db10(double v) { return 10.0 * log10(fmax(v, kEps)); }

// block means of `nwin` windows of `w` samples, in dB, into sh.win (nwin <= LB_MAXWIN): one warp per window
__device__ void fill_windows(const SqView& sq, int nwin, int64_t w, LundebyShared& sh) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < nwin; k += LB_THREADS / 32) {
    const int64_t a = (int64_t)k * w;
    double acc = range_partial(sq, a, a + w, lane, 32);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sh.win[k] = db10(acc / (double)w);
  }
  __syncthreads();
}

struct Regress { double slope, icpt; };
__device__ Regress linregress_dev(const double* y, double wd, int s, int e) {
  // x[k] = k*wd + wd/2 ; scipy: slope = cov(x,y)/var(x) (bias=1), intercept = ymean - slope*xmean
  const int n = e - s;
  double xm = 0.0, ym = 0.0;
  for (int k = s; k < e; ++k) { xm += (double)k * wd + wd / 2; ym += y[k]; }
  xm /= n; ym /= n;
  double sxx = 0.0, sxy = 0.0;
  for (int k = s; k < e; ++k) {
    const double dx = ((double)k * wd + wd / 2) - xm;
    sxx += dx * dx;
    sxy += dx * (y[k] - ym);
  }
  Regress r;
  r.slope = (sxy / n) / (sxx / n);
  r.icpt = ym - r.slope * xm;
  return r;
}

__device__ __forceinline__ int first_le(const double* w, int n, double thr) {
  for (int k = 0; k < n; ++k)
    if (w[k] <= thr) return k;
  return -1;
}

// argmin_i |i*step - t| over i in [0, ns), first index on ties (np.argmin(np.abs(t_squared - t)))
__device__ int64_t nearest_t(double t, double step, int64_t ns) {
  if (ns <= 1 || !(step > 0.0)) return 0;
  int64_t i0 = (int64_t)floor(t / step);
  i0 = max((int64_t)0, min(i0, ns - 1));
  int64_t best = i0;
  double bd = fabs((double)i0 * step - t);
  for (int64_t c = max((int64_t)0, i0 - 2); c <= min(ns - 1, i0 + 2); ++c) {
    const double d = fabs((double)c * step - t);
    if (d < bd || (d == bd && c < best)) { bd = d; best = c; }
  }
  return best;
}

// ---- grid-wide passes feeding k_lundeby: the two full reads of the analysis window [pk, pk + 2 s) ----------------
__device__ __forceinline__ int lb_block_size(int64_t ns) {
  int bs = 32;
  while ((ns + bs - 1) / bs > LB_MAXBLK) bs <<= 1;
  return bs;
}

// max |v[pk + i]|, i < ns, grid = (rows, chunks)
__global__ void __launch_bounds__(256) k_lb_segmax(const double* __restrict__ x, const ViewArgs va, int64_t r0, int fs,
                                                   const int64_t* __restrict__ peaks,
                                                   unsigned long long* __restrict__ segmax) {
  const int64_t r = blockIdx.x;
  const RowView ir(x, va, r0 + r);
  if (ir.n < 10) return;
  const int64_t pk = peaks[r];
  const int64_t ns = min(pk + (int64_t)(2 * fs), ir.n) - pk;
  const int64_t jh = ir.hinted(va, r0 + r);
  if (jh >= pk && jh < pk + ns) {  // the known maximum lies inside the analysis window
    if (blockIdx.y == 0 && threadIdx.x == 0)
      segmax[r] = (unsigned long long)__double_as_longlong(va.hint_max[r0 + r]);
    return;
  }
  int64_t lo, hi;
  chunk_bounds(ns, &lo, &hi);
  if (lo >= hi) return;
  double m = 0.0;
  int64_t i = lo + threadIdx.x;
  for (; i + 768 < hi; i += 1024) {
    const double a0 = ir[pk + i], a1 = ir[pk + i + 256], a2 = ir[pk + i + 512], a3 = ir[pk + i + 768];
    m = fmax(fmax(m, fmax(fabs(a0), fabs(a1))), fmax(fabs(a2), fabs(a3)));
  }
  for (; i < hi; i += 256) m = fmax(m, fabs(ir[pk + i]));
  for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(segmax + r, (unsigned long long)__double_as_longlong(m));
}

// blk[r][k] = sum of sq over the k-th aligned block of bs samples (one warp per block), grid = (rows, chunks)
__global__ void __launch_bounds__(256) k_lb_blocksums(const double* __restrict__ x, const ViewArgs va, int64_t r0, int fs,
                                                      const int64_t* __restrict__ peaks,
                                                      const unsigned long long* __restrict__ segmax,
                                                      double* __restrict__ blk) {
  const int64_t r = blockIdx.x;
  const RowView ir(x, va, r0 + r);
  if (ir.n < 10) return;
  const int64_t pk = peaks[r];
  const int64_t ns = min(pk + (int64_t)(2 * fs), ir.n) - pk;
  const int bs = lb_block_size(ns);
  const int nblk = (int)((ns + bs - 1) / bs);
  const double mx = __longlong_as_double((long long)segmax[r]);
  const SqView sq{&ir, pk, mx, mx >= kEps, nullptr, bs, ExactDivisor(mx)};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (nblk + gridDim.y - 1) / gridDim.y;
  const int k_end = min(nblk, (int)(blockIdx.y + 1) * per);
  // segment sample i is ir[pk + i] = ir.p[pk + i - sh] with weight 1 when lo <= pk + i < hi and pk + i - sh >= fade_len
  const int ns32 = (int)ns;  // ns <= 2 fs
  const int in_lo = (int)min(max(max(ir.lo - pk, ir.fade_len + ir.sh - pk), (int64_t)0), ns);
  const int in_hi = (int)max(min(ir.hi - pk, ns), (int64_t)0);
  const double* direct = ir.p + (pk - ir.sh);
  for (int k0 = blockIdx.y * per + warp; k0 < k_end; k0 += 64) {  // 8 warps x 8 blocks in flight
    double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    if (bs == 32 && k0 * 32 >= in_lo && min(ns32, (k0 + 57) * 32) <= in_hi) {
      // one sample per lane and block, all of them inside the unfaded valid part of the view: eight predicated
      // loads issued back to back (no control flow between them), then the arithmetic.  A lane without a sample
      // contributes +0.0, which is what its untouched accumulator held before.
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + u * 8, i = k * 32 + lane;
        v[u] = (k < k_end && i < ns32) ? __ldg(direct + i) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = sq.of(v[u]);
    } else {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + u * 8;
        if (k < k_end) {
          const int64_t a = (int64_t)k * bs, b = min(ns, a + bs);
          for (int64_t i = a + lane; i < b; i += 32) acc[u] += sq(i);
        }
      }
    }
    // eight xor-butterfly reductions packed into one: after the offset-16 step a lane only carries four of the eight
    // sums, after offset 8 two, after offset 4 one (9 double shuffles instead of 40).  Every partial sum has the
    // same two operands as in the plain butterfly, so the results are bit-identical; lane 4u ends up with block u.
    const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4;
    double x4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      x4[i] = (hi16 ? acc[4 + i] : acc[i]) + __shfl_xor_sync(0xffffffffu, hi16 ? acc[i] : acc[4 + i], 16);
    double y2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
      y2[i] = (hi8 ? x4[2 + i] : x4[i]) + __shfl_xor_sync(0xffffffffu, hi8 ? x4[i] : x4[2 + i], 8);
    double y = (hi4 ? y2[1] : y2[0]) + __shfl_xor_sync(0xffffffffu, hi4 ? y2[0] : y2[1], 4);
    for (int o = 2; o; o >>= 1) y += __shfl_xor_sync(0xffffffffu, y, o);
    const int k = k0 + (lane >> 2) * 8;
    if ((lane & 3) == 0 && k < k_end) blk[(size_t)r * LB_MAXBLK + k] = y;
  }
}

__global__ void __launch_bounds__(LB_THREADS) k_lundeby(const double* __restrict__ x, const ViewArgs va, int64_t r0, int fs,
                                                        const int64_t* __restrict__ peaks,
                                                        const unsigned long long* __restrict__ segmax,
                                                        const double* __restrict__ blk_g, int64_t* __restrict__ peak_out,
                                                        int64_t* __restrict__ knee_out, double* __restrict__ noise_out,
                                                        int64_t* __restrict__ win_out) {
  extern __shared__ __align__(16) unsigned char lb_smem[];
  LundebyShared& sh = *reinterpret_cast<LundebyShared*>(lb_smem);
  const int r = blockIdx.x;
  const RowView ir(x, va, r0 + r);
  const int64_t n_ir = ir.n;
  const bool lead = threadIdx.x == 0;
#define LB_RETURN(pk_, knee_, noise_, w_)                                                    \
  do {                                                                                       \
    if (lead) { peak_out[r] = (pk_); knee_out[r] = (knee_); noise_out[r] = (noise_); win_out[r] = (w_); } \
    return;                                                                                  \
  } while (0)

  if (n_ir < 10) LB_RETURN(0, n_ir, -200.0, n_ir > 0 ? n_ir : 1);
  const int64_t pk = peaks[r];
  const int64_t stop = min(pk + (int64_t)(2 * fs), n_ir);
  const int64_t ns = stop - pk;  // >= 1 because pk <= n_ir - 1
  // the two full reads of the window (max, block sums of sq) were done by grid-wide kernels
  const double mx = __longlong_as_double((long long)segmax[r]);
  SqView sq{&ir, pk, mx, mx >= kEps, sh.blk, lb_block_size(ns), ExactDivisor(mx)};
  {
    const int nblk = (int)((ns + sq.bs - 1) / sq.bs);
    const double* mine = blk_g + (size_t)r * LB_MAXBLK;
    for (int k = threadIdx.x; k < nblk; k += LB_THREADS) sh.blk[k] = mine[k];
    __syncthreads();
  }
  const double step = ns > 1 ? ((double)ns / (double)fs) / (double)(ns - 1) : 0.0;  // linspace(0, ns/fs, ns)
  const double t_last = (double)(ns - 1) * step;

  // ---- pass 1: 30 ms windows ---------------------------------------------------------------
  double wd = 0.03;
  int64_t nwin = (int64_t)((double)ns / (double)fs / wd);
  if (nwin == 0) {
    const double nf = 10.0 * log10(fmax(range_mean(sq, 0, ns, sh), kEps));
    LB_RETURN(pk, pk + ns, nf, max((int64_t)1, ns));
  }
  int64_t w = max((int64_t)((double)ns / (double)nwin), (int64_t)1);
  const int64_t w_first = w;
  if (nwin > LB_MAXWIN) nwin = LB_MAXWIN;  // cannot happen for 30 ms windows of <= 2 s
  fill_windows(sq, (int)nwin, w, sh);

  int64_t tail0 = (int64_t)((double)ns * 0.9);
  if (tail0 >= ns) tail0 = 0;
  const double noise = 10.0 * log10(fmax(range_mean(sq, tail0, ns, sh), kEps));

  // scalar section (thread 0), results broadcast through shared memory
  if (lead) {
    int status = 0;  // 0 continue, 1 return full-length
    int slope_end = (int)nwin;
    const int c0 = first_le(sh.win, (int)nwin, noise + 10.0);
    if (c0 > 0) slope_end = c0;
    double slope = 0.0, icpt = 0.0;
    if (slope_end < 2) {
      if (nwin >= 2) slope_end = (int)nwin; else status = 1;
    }
    if (!status) {
      const Regress rg = linregress_dev(sh.win, wd, 0, slope_end);
      slope = rg.slope; icpt = rg.icpt;
      if (isnan(slope) || fabs(slope) < kEps) status = 1;
    }
    sh.ibcast[0] = status;
    sh.bcast[0] = slope;
    sh.bcast[1] = icpt;
  }
  __syncthreads();
  if (sh.ibcast[0] == 1) LB_RETURN(pk, pk + ns, noise, w_first);
  const double slope = sh.bcast[0], icpt = sh.bcast[1];
  double knee_t = fmin(fmax((noise - icpt) / slope, 0.0), t_last);
  __syncthreads();

  // ---- pass 2: three windows per 10 dB of decay ------------------------------------------------
  const double den = fabs(slope) * 3;
  wd = den < kEps ? t_last / 3.0 : 10.0 / den;
  int64_t n2 = wd <= kEps ? 1 : (int64_t)((double)ns / (double)fs / wd);
  if (n2 == 0) n2 = 1;
  w = max((int64_t)((double)ns / (double)n2), (int64_t)1);
  const int64_t n_t = n2;              // length of t_windows (kept even if the windows array is shortened)
  if (n2 * w > ns) n2 = ns / w;
  // n2 <= 4000 < LB_MAXWIN always: the first-pass regression runs over window levels clamped to [-200 dB, 0 dB]
  // (EPSILON = 1e-20 floor, peak normalised to 1) at 30 ms spacing, and an OLS slope over >= 2 such points is at most
  // 200 dB / 30 ms, so wd >= 10 / (3 * 6667 dB/s) = 0.5 ms and n2 = (ns / fs) / wd <= 2 s / 0.5 ms
  // (tests: test_lundeby_steepest_possible_decay drives the reference to 3840 windows).  Not a tolerance: an
  // assertion — a violated bound stops the launch with a device error instead of returning a different integer.
  if (n2 > LB_MAXWIN) __trap();
  fill_windows(sq, (int)n2, w, sh);
  const int nw = (int)n2;
  const double t_win_last = (double)(n_t - 1) * wd + wd / 2;

  // knee index in window units
  if (lead) {
    int k_idx = -1;
    for (int64_t k = 0; k < n_t; ++k)
      if ((double)k * wd + wd / 2 >= knee_t) { k_idx = (int)k; break; }
    double kt = knee_t, kv;
    if (k_idx < 0 || k_idx >= nw) {
      kt = t_win_last;
      k_idx = (int)(n_t - 1);
      kv = sh.win[nw - 1];
    } else {
      kv = sh.win[k_idx];
    }
    sh.bcast[0] = kt;
    sh.bcast[1] = kv;
    sh.ibcast[0] = k_idx;
  }
  __syncthreads();
  double knee_t_it = sh.bcast[0], knee_val_it = sh.bcast[1];
  long long k_idx_it = sh.ibcast[0];
  double noise_it = noise;
  __syncthreads();

  for (int iter = 0; iter < 5; ++iter) {
    // -- noise segment bounds (thread 0) --
    if (lead) {
      long long go = 1;
      int64_t a = 0, b = 0;
      const int nf0 = first_le(sh.win, nw, knee_val_it - 5);
      if (nf0 < 0) go = 0;
      if (go) {
        const double t0 = fmax((double)nf0 * wd + wd / 2, 0.1 * t_last);
        if (t0 > t_win_last) go = 0;
        if (go) {
          const double t1 = fmin(t0 + knee_t_it, t_last);
          a = nearest_t(t0, step, ns);
          b = nearest_t(t1, step, ns);
          if (a >= b) go = 0;
        }
      }
      sh.ibcast[0] = go; sh.ibcast[1] = a; sh.ibcast[2] = b;
    }
    __syncthreads();
    const long long go = sh.ibcast[0];
    const int64_t a = sh.ibcast[1], b = sh.ibcast[2];
    __syncthreads();
    if (!go) break;
    noise_it = 10.0 * log10(fmax(range_mean(sq, a, b, sh), kEps));
    // -- late slope (thread 0) --
    if (lead) {
      long long state = 1;  // 1 continue iterating, 0 break, 2 converged (break after update)
      int e = first_le(sh.win, nw, noise_it + 8);
      int s = first_le(sh.win, nw, noise_it + 8 + 20);
      double nkt = knee_t_it, nkv = knee_val_it;
      long long nki = k_idx_it;
      if (e < 0 || s < 0) state = 0;
      if (state) {
        e -= 1;
        s = max(s - 1, 0);
        if (e <= s + 1) state = 0;
      }
      if (state) {
        const Regress rg = linregress_dev(sh.win, wd, s, e);
        if (isnan(rg.slope) || fabs(rg.slope) < kEps) {
          state = 0;
        } else {
          const double t_first = wd / 2;
          const double new_t = fmin(fmax((noise_it - rg.icpt) / rg.slope, t_first), t_win_last);
          long long new_idx = -1;
          for (int64_t k = 0; k < n_t; ++k)
            if ((double)k * wd + wd / 2 >= new_t) { new_idx = k; break; }
          if (new_idx < 0) new_idx = n_t - 1;
          if (new_idx == k_idx_it) {
            nkt = (double)k_idx_it * wd + wd / 2;
            state = 2;
          } else {
            nki = new_idx;
            nkt = (double)new_idx * wd + wd / 2;
            nkv = new_idx < nw ? sh.win[new_idx] : sh.win[nw - 1];
          }
        }
      }
      sh.ibcast[0] = state; sh.ibcast[1] = nki;
      sh.bcast[0] = nkt; sh.bcast[1] = nkv;
    }
    __syncthreads();
    const long long state = sh.ibcast[0];
    k_idx_it = sh.ibcast[1];
    knee_t_it = sh.bcast[0];
    knee_val_it = sh.bcast[1];
    __syncthreads();
    if (state != 1) break;
  }
  LB_RETURN(pk, pk + nearest_t(knee_t_it, step, ns), noise_it, w);
#undef LB_RETURN
}

// ------------------------------------------------------------------------------------------------
// onset cross-correlation: one CTA per pair, direct evaluation in FP64
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_xcorr(const double* __restrict__ x, const int64_t* __restrict__ a_off,
                                               const int64_t* __restrict__ a_len, const int64_t* __restrict__ b_off,
                                               const int64_t* __restrict__ b_len, const double* __restrict__ fade,
                                               int64_t fade_len, int64_t* __restrict__ lag_out) {
  extern __shared__ double smx[];
  const int r = blockIdx.x;
  const int na = (int)a_len[r], nb = (int)b_len[r];
  // Sample i lives at slot i + i/8: consecutive threads own lags XR = 8 apart, so at every step their a / b indices are
  // 8 apart too -- unskewed that is a 64-byte stride, a 16-way bank conflict on every load (ncu: LSU wavefronts 82 % of
  // peak, FP64 pipe 26 %); with the skew the stride is 9 doubles and a warp's load touches every bank once.
  auto slot = [](int i) { return i + (i >> 3); };
  double* a = smx;
  double* b = smx + slot(na) + 1;
  // segments start at the head of a cropped response: the fade-in of crop_heads applies to their first samples
  for (int i = threadIdx.x; i < na; i += blockDim.x) {
    double v = x[a_off[r] + i];
    a[slot(i)] = (fade && i < fade_len) ? __dmul_rn(v, fade[i]) : v;
  }
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    double v = x[b_off[r] + i];
    b[slot(i)] = (fade && i < fade_len) ? __dmul_rn(v, fade[i]) : v;
  }
  __syncthreads();
  // z[k] = sum_l a[l] * b[l - k + nb - 1],  k in [0, na + nb - 1).  Every thread owns XR consecutive lags and slides a
  // register window of b along l: one broadcast load of a[l] and one load of b feed XR FMAs (evaluated lag by lag the
  // kernel was bound by shared-memory bandwidth: 3 wavefronts per warp-FMA).  Each lag is still ONE chain summed in
  // increasing l (out-of-range taps contribute exact zeros), so the values are those of the plain loop.
  constexpr int XR = 8;
  double best = -DBL_MAX;
  int best_k = 0x7fffffff;
  const int nk = na + nb - 1;
  auto b_at = [&](int i) -> double { return (i >= 0 && i < nb) ? b[slot(i)] : 0.0; };
  for (int k0 = threadIdx.x * XR; k0 < nk; k0 += blockDim.x * XR) {
    const int shift0 = k0 - (nb - 1);  // lag k0 + r pairs a[l] with b[l - shift0 - r]
    const int lo = max(0, shift0), hi = min(na, nb + shift0 + XR - 1);
    double acc[XR], w[XR];
#pragma unroll
    for (int r = 0; r < XR; ++r) { acc[r] = 0.0; w[r] = b_at(lo - shift0 - r); }
    int l = lo;
    for (; l + XR <= hi; l += XR) {
#pragma unroll
      for (int u = 0; u < XR; ++u) {
        // at step l + u the window slot of lag r is w[(r - u) mod XR]
        const double av = a[slot(l + u)];
#pragma unroll
        for (int r = 0; r < XR; ++r) acc[r] = fma(av, w[(r - u + XR) % XR], acc[r]);
        // slot (XR - 1 - u + XR) % XR ... the oldest value (lag XR-1) is replaced by the next b for lag 0
        w[(XR - 1 - u + XR) % XR] = b_at(l + u + 1 - shift0);
      }
    }
    for (; l < hi; ++l) {  // tail: plain rotation
      const double av = a[slot(l)];
#pragma unroll
      for (int r = 0; r < XR; ++r) acc[r] = fma(av, w[r], acc[r]);
#pragma unroll
      for (int r = XR - 1; r > 0; --r) w[r] = w[r - 1];
      w[0] = b_at(l + 1 - shift0);
    }
#pragma unroll
    for (int r = 0; r < XR; ++r)
      if (k0 + r < nk && acc[r] > best) { best = acc[r]; best_k = k0 + r; }
  }
  // block arg-max, first index on ties
  __shared__ double sv[256];
  __shared__ int sk[256];
  sv[threadIdx.x] = best;
  sk[threadIdx.x] = best_k;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) {
      const double v2 = sv[threadIdx.x + o];
      const int k2 = sk[threadIdx.x + o];
      if (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && k2 < sk[threadIdx.x])) {
        sv[threadIdx.x] = v2;
        sk[threadIdx.x] = k2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) lag_out[r] = (nk > 0 && na > 0) ? (int64_t)sk[0] - (int64_t)(na - 1) : 0;
}

// ------------------------------------------------------------------------------------------------
// fused gather / shift / fade / gain
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gather_window(const double* __restrict__ src, const int64_t* __restrict__ src_start,
                                                       const int64_t* __restrict__ shift, const int64_t* __restrict__ valid_lo,
                                                       const int64_t* __restrict__ valid_hi, int64_t out_len,
                                                       const double* __restrict__ fade_in, int64_t fade_in_len,
                                                       const double* __restrict__ fade_out, int64_t fade_out_len, double gain,
                                                       int apply_gain, double* __restrict__ dst, int64_t ld_dst,
                                                       const int64_t* __restrict__ out_lens,
                                                       const int64_t* __restrict__ dst_offsets) {
  const int r = blockIdx.y;
  if (out_lens) out_len = out_lens[r];  // ragged form: per-row length and destination offset
  if (out_len < fade_out_len) fade_out_len = 0;
  if (dst_offsets) { dst += dst_offsets[r]; ld_dst = 0; }
  const int64_t s0 = src_start[r], sh = shift ? shift[r] : 0;
  const int64_t lo = valid_lo ? valid_lo[r] : 0, hi = valid_hi ? valid_hi[r] : out_len;
  const int64_t fo0 = out_len - fade_out_len;
  // a block walks spans of 4 x 256 samples; the four source loads of a thread are issued before any of the
  // dependent fade / gain arithmetic and stores (no control flow between them)
  const double* from = src + (s0 - sh);
  for (int64_t base = blockIdx.x * (int64_t)1024; base < out_len; base += (int64_t)gridDim.x * 1024) {
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = base + u * 256 + threadIdx.x;
      v[u] = (j < out_len && j >= lo && j < hi) ? from[j] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = base + u * 256 + threadIdx.x;
      if (j >= out_len) break;
      double w = v[u];
      const int64_t m = j - sh;  // index into the head-cropped array
      if (j >= lo && j < hi && m < fade_in_len) w = __dmul_rn(w, fade_in[m]);
      if (j >= fo0 && fade_out_len > 0) w = __dmul_rn(w, fade_out[j - fo0]);
      if (apply_gain) w = __dmul_rn(w, gain);
      dst[(size_t)r * ld_dst + j] = w;
    }
  }
}

}  // namespace

extern "C" {

static ViewArgs make_view(const int64_t* offsets, const int64_t* lengths, const impb_view* v) {
  ViewArgs a{offsets, lengths, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, 0, 0};
  if (v) {
    // block maxima describe untouched rows: ignored for shifted / faded views
    if (v->blk_max && v->nblk > 0 && v->blk_len > 0 && !v->shift && !(v->fade_in_len > 0) && !v->valid_lo && !v->valid_hi) {
      a.blk_max = v->blk_max; a.nblk = v->nblk; a.blk_len = v->blk_len; a.blk_first = v->blk_first;
    }
    if (v->hint_max && v->hint_argmax) { a.hint_max = v->hint_max; a.hint_arg = v->hint_argmax; }
    a.shift = v->shift;
    a.lo = v->valid_lo;
    a.hi = v->valid_hi;
    a.fade = v->fade_in_len > 0 ? v->fade_in : nullptr;
    a.fade_len = v->fade_in_len;
  }
  return a;
}

int32_t impb_peak_index(impb_ctx* ctx, const double* x, const int64_t* offsets, const int64_t* lengths,
                        const impb_view* view, const double* known_max, double height, int64_t* peak_out,
                        int64_t batch, int64_t max_len, void* stream) {
  if (!ctx || !x || !offsets || !lengths || !peak_out || batch < 0 || !(height > 0.0)) {
    set_error("impb_peak_index: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  return run_peak_index(ctx, x, make_view(offsets, lengths, view), height, peak_out, batch,
                        std::max<int64_t>(max_len, 1), (cudaStream_t)stream, known_max);
}

int32_t impb_lundeby(impb_ctx* ctx, const double* x, const int64_t* offsets, const int64_t* lengths,
                     const impb_view* view, int32_t fs, int64_t* peak_out, int64_t* knee_out, double* noise_db_out,
                     int64_t* win_out, int64_t batch, int64_t max_len, void* stream) {
  if (!ctx || !x || !offsets || !lengths || !peak_out || !knee_out || !noise_db_out || !win_out || batch < 0 ||
      fs <= 0) {
    set_error("impb_lundeby: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t* peaks = nullptr;
  IMPB_CUDA(cudaMallocAsync(&peaks, sizeof(int64_t) * batch, st));
  const ViewArgs va = make_view(offsets, lengths, view);
  int rc = run_peak_index(ctx, x, va, 0.12589, peaks, batch, std::max<int64_t>(max_len, 1), st);
  if (!rc) {
    cudaError_t e = cudaFuncSetAttribute(k_lundeby, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(LundebyShared));
    const int64_t rows_per = std::min<int64_t>(batch, 16384);  // 24 KB of block sums per row
    unsigned long long* segmax = nullptr;
    double* blk = nullptr;
    if (e == cudaSuccess) e = cudaMallocAsync(&segmax, sizeof(unsigned long long) * rows_per, st);
    if (e == cudaSuccess) e = cudaMallocAsync(&blk, sizeof(double) * LB_MAXBLK * rows_per, st);
    // CTAs per row of the two streaming passes: 4096 samples each when few rows must fill the GPU (one recording set),
    // 16 384 when the rows alone do (a CTA of 4096 samples lives ~5 us, a third of it prologue: 43 008 CTAs per step)
    const int64_t fine = std::max<int64_t>(1, std::min<int64_t>((2 * (int64_t)fs + 4095) / 4096, 256));
    const int64_t coarse = std::max<int64_t>(1, (2 * (int64_t)fs + 16383) / 16384);
    for (int64_t b0 = 0; b0 < batch && e == cudaSuccess; b0 += rows_per) {
      const int64_t nb = std::min<int64_t>(rows_per, batch - b0);
      const unsigned chunks = (unsigned)std::min<int64_t>(fine, std::max<int64_t>(coarse, (1184 + nb - 1) / nb));
      e = cudaMemsetAsync(segmax, 0, sizeof(unsigned long long) * nb, st);
      k_lb_segmax<<<dim3((unsigned)nb, chunks), 256, 0, st>>>(x, va, b0, fs, peaks + b0, segmax);
      k_lb_blocksums<<<dim3((unsigned)nb, chunks), 256, 0, st>>>(x, va, b0, fs, peaks + b0, segmax, blk);
      k_lundeby<<<(unsigned)nb, LB_THREADS, sizeof(LundebyShared), st>>>(x, va, b0, fs, peaks + b0, segmax, blk,
                                                                        peak_out + b0, knee_out + b0, noise_db_out + b0,
                                                                        win_out + b0);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (segmax) cudaFreeAsync(segmax, st);
    if (blk) cudaFreeAsync(blk, st);
    if (e != cudaSuccess) rc = cuda_fail(e, "k_lundeby");
  }
  cudaFreeAsync(peaks, st);
  return rc;
}

int32_t impb_xcorr_lag(impb_ctx* ctx, const double* x, const int64_t* a_offsets, const int64_t* a_lengths,
                       const int64_t* b_offsets, const int64_t* b_lengths, const double* fade_in, int64_t fade_in_len,
                       int64_t* lag_out, int64_t batch, int64_t max_seg_len, void* stream) {
  if (!ctx || !x || !a_offsets || !a_lengths || !b_offsets || !b_lengths || !lag_out || batch < 0 || max_seg_len < 0) {
    set_error("impb_xcorr_lag: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0) return IMPB_OK;
  const size_t smem = 2 * ((size_t)max_seg_len + (size_t)max_seg_len / 8 + 2) * sizeof(double);  // skewed slots
  if (smem > ctx->smem_optin) {
    set_error("impb_xcorr_lag: segments of %lld samples do not fit shared memory", (long long)max_seg_len);
    return IMPB_EUNSUPPORTED;
  }
  DeviceGuard g(ctx->device);
  IMPB_CUDA(cudaFuncSetAttribute(k_xcorr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_xcorr<<<(unsigned)batch, 256, smem, (cudaStream_t)stream>>>(x, a_offsets, a_lengths, b_offsets, b_lengths,
                                                                fade_in_len > 0 ? fade_in : nullptr, fade_in_len, lag_out);
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

int32_t impb_gather_window(impb_ctx* ctx, const double* src, const int64_t* src_start, const int64_t* shift,
                           const int64_t* valid_lo, const int64_t* valid_hi, int64_t out_len, const double* fade_in,
                           int64_t fade_in_len, const double* fade_out, int64_t fade_out_len, double gain, double* dst,
                           int64_t ld_dst, int64_t batch, void* stream) {
  if (!ctx || !src || !src_start || !dst || batch < 0 || out_len < 0 || ld_dst < out_len || fade_in_len < 0 ||
      fade_out_len < 0 || fade_out_len > out_len || (fade_in_len > 0 && !fade_in) || (fade_out_len > 0 && !fade_out)) {
    set_error("impb_gather_window: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0 || out_len == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  const int thr = 256;
  int64_t bx = std::max<int64_t>(1, std::min<int64_t>((out_len + thr * 4 - 1) / (thr * 4), 1024));
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = (int)std::min<int64_t>(65535, batch - b0);
    k_gather_window<<<dim3((unsigned)bx, nb), thr, 0, (cudaStream_t)stream>>>(
        src, src_start + b0, shift ? shift + b0 : nullptr, valid_lo ? valid_lo + b0 : nullptr,
        valid_hi ? valid_hi + b0 : nullptr, out_len, fade_in, fade_in_len, fade_out, fade_out_len, gain,
        gain != 1.0 ? 1 : 0, dst + (size_t)b0 * ld_dst, ld_dst, nullptr, nullptr);
  }
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

int32_t impb_gather_rows(impb_ctx* ctx, const double* src, const int64_t* src_start, const int64_t* shift,
                         const int64_t* valid_lo, const int64_t* valid_hi, const int64_t* out_lens,
                         const int64_t* dst_offsets, int64_t max_out_len, const double* fade_in, int64_t fade_in_len,
                         const double* fade_out, int64_t fade_out_len, double gain, double* dst, int64_t batch,
                         void* stream) {
  if (!ctx || !src || !src_start || !out_lens || !dst_offsets || !dst || batch < 0 || max_out_len < 0 || fade_in_len < 0 ||
      fade_out_len < 0 || (fade_in_len > 0 && !fade_in) || (fade_out_len > 0 && !fade_out)) {
    set_error("impb_gather_rows: bad argument");
    return IMPB_EINVAL;
  }
  if (batch == 0 || max_out_len == 0) return IMPB_OK;
  DeviceGuard g(ctx->device);
  const int thr = 256;
  int64_t bx = std::max<int64_t>(1, std::min<int64_t>((max_out_len + thr * 4 - 1) / (thr * 4), 1024));
  for (int64_t b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = (int)std::min<int64_t>(65535, batch - b0);
    k_gather_window<<<dim3((unsigned)bx, nb), thr, 0, (cudaStream_t)stream>>>(
        src, src_start + b0, shift ? shift + b0 : nullptr, valid_lo ? valid_lo + b0 : nullptr,
        valid_hi ? valid_hi + b0 : nullptr, 0, fade_in, fade_in_len, fade_out, fade_out_len, gain, gain != 1.0 ? 1 : 0,
        dst, 0, out_lens + b0, dst_offsets + b0);
  }
  IMPB_CUDA(cudaGetLastError());
  return IMPB_OK;
}

}  // extern "C"
