// Context, error reporting and the FFT planner (host code).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace impb {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return IMPB_ECUDA;
}

static const int kRadices[] = {16, 15, 12, 10, 9, 8, 6, 5, 4, 3, 2};

static double radix_cost(int r) { return r * (3.0 + log2((double)r)); }

static void dfs(int rem, int max_r, std::vector<int>& cur, int S, int tile, int threads, double* best_cost,
                std::vector<int>* best) {
  if (rem == 1) {
    double cost = 0;
    for (int r : cur) {
      const int items = (S / r) * tile;
      cost += (double)((items + threads - 1) / threads) * radix_cost(r);
    }
    cost += 4.0 * cur.size();  // barrier + loop overhead per stage
    if (cost < *best_cost) {
      *best_cost = cost;
      *best = cur;
    }
    return;
  }
  if ((int)cur.size() >= MAX_STAGES) return;
  for (int r : kRadices) {
    if (r > max_r || rem % r) continue;
    cur.push_back(r);
    dfs(rem / r, r, cur, S, tile, threads, best_cost, best);
    cur.pop_back();
  }
}

static double subplan_cost(int S, int tile, int threads, std::vector<int>* radices) {
  double best = 1e300;
  std::vector<int> cur, out;
  dfs(S, 16, cur, S, tile, threads, &best, &out);
  if (radices) *radices = out;
  return best;
}

bool fixed_radices(int S, int* r) {
  // sub-FFT lengths with compile-time specialised kernels (fft_fixed.cuh); keep in sync with the dispatch switches
  // in fft2d.cuh (every length: plain forward / inverse passes) and deconv.cu (768: the fused K1 triple)
  switch (S) {
    case 1600: r[0] = 16; r[1] = 10; r[2] = 10; return true;  // 192 kHz K1 plan (1536 x 1600), half-width tiles only
    case 1536: r[0] = 12; r[1] = 8; r[2] = 16; return true;
    case 768: r[0] = 8; r[1] = 12; r[2] = 8; return true;
    case 675: r[0] = 15; r[1] = 9; r[2] = 5; return true;
    case 432: r[0] = 12; r[1] = 9; r[2] = 4; return true;
    case 400: r[0] = 10; r[1] = 8; r[2] = 5; return true;
    case 300: r[0] = 12; r[1] = 5; r[2] = 5; return true;
    case 256: r[0] = 8; r[1] = 8; r[2] = 4; return true;
    case 243: r[0] = 9; r[1] = 9; r[2] = 3; return true;
    case 162: r[0] = 9; r[1] = 6; r[2] = 3; return true;
    case 160: r[0] = 8; r[1] = 5; r[2] = 4; return true;
    case 135: r[0] = 5; r[1] = 9; r[2] = 3; return true;
    case 128: r[0] = 8; r[1] = 4; r[2] = 4; return true;
    case 125: r[0] = 5; r[1] = 5; r[2] = 5; return true;
    case 120: r[0] = 8; r[1] = 5; r[2] = 3; return true;
    case 108: r[0] = 4; r[1] = 9; r[2] = 3; return true;
    case 81: r[0] = 3; r[1] = 9; r[2] = 3; return true;
    case 75: r[0] = 5; r[1] = 5; r[2] = 3; return true;
    default: return false;
  }
}

bool fixed_tile_ok(int S, bool full_tile) {
  // which tile width the specialised kernels of a length are built for: 1536 / 1600 (88.2 ... 192 kHz K1 plans) only
  // exist for half-width tiles, 768 for both (it pairs with 1536 at 88.2 / 96 kHz), everything else for full width
  return S > 1024 ? !full_tile : (full_tile || S == 768);
}

std::vector<double> stage_table_full(const SubPlan& sp, int s, bool transposed) {
  // every stage twiddle exp(-2 pi i np p / len_s), p = 1..R-1, at [np*(R-1) + p-1] (column pass) or
  // [(p-1)*stride + np] (row pass): the exact-twiddle tables of the plain specialised passes
  const int R = sp.radix[s], st = sp.stride[s], len = sp.len[s];
  std::vector<double> t(2 * (size_t)st * (R - 1));
  for (int np = 0; np < st; ++np)
    for (int p = 1; p < R; ++p) {
      double re, im;
      unit_root((int64_t)np * p, len, &re, &im);
      const size_t i = transposed ? (size_t)(p - 1) * st + np : (size_t)np * (R - 1) + (p - 1);
      t[2 * i] = re;
      t[2 * i + 1] = im;
    }
  return t;
}

bool make_subplan(int S, int tile, int threads, SubPlan* sp) {
  if (S < 2) return false;
  std::vector<int> rad;
  int fr[3];
  if (fixed_radices(S, fr)) {
    rad.assign(fr, fr + 3);
  } else if (subplan_cost(S, tile, threads, &rad) >= 1e299 || rad.empty()) {
    return false;
  }
  memset(sp, 0, sizeof(*sp));
  sp->S = S;
  sp->nst = (int)rad.size();
  sp->qmode = (S % 4 == 0) ? 1 : 0;
  sp->ntab = sp->qmode ? S / 4 : S;
  int len = S;
  for (int s = 0; s < sp->nst; ++s) {
    sp->radix[s] = rad[s];
    sp->len[s] = len;
    sp->stride[s] = len / rad[s];
    sp->mult[s] = S / len;
    len /= rad[s];
  }
  return true;
}

std::vector<int> subplan_perm(const SubPlan& sp) {
  std::vector<int> perm(sp.S);
  for (int pos = 0; pos < sp.S; ++pos) {
    int rem = pos, k = 0, w = 1;
    for (int s = 0; s < sp.nst; ++s) {
      const int p = rem / sp.stride[s];
      rem -= p * sp.stride[s];
      k += p * w;
      w *= sp.radix[s];
    }
    perm[pos] = k;
  }
  return perm;
}

void unit_root(int64_t k, int64_t N, double* re, double* im) {
  k %= N;
  if (k < 0) k += N;
  // fold to the first octant so that the long double angle stays small and exact symmetries hold
  const long double two_pi = 6.283185307179586476925286766559005768L;
  int64_t k8 = k * 8;
  int oct = (int)(k8 / N);          // 0..7
  int64_t r = k * 8 - (int64_t)oct * N;  // remainder in units of 1/(8N) turns ... angle = 2pi*(oct/8 + r/(8N))
  long double phi = two_pi * ((long double)r / (long double)(8 * N));
  long double c = cosl(phi), s = sinl(phi);  // phi in [0, pi/4)
  // rotate by oct * 45 degrees
  const long double h = 0.707106781186547524400844362104849039L;
  long double cr, sr;
  switch (oct) {
    case 0: cr = c; sr = s; break;
    case 1: cr = h * (c - s); sr = h * (c + s); break;
    case 2: cr = -s; sr = c; break;
    case 3: cr = -h * (c + s); sr = h * (c - s); break;
    case 4: cr = -c; sr = -s; break;
    case 5: cr = -h * (c - s); sr = -h * (c + s); break;
    case 6: cr = s; sr = -c; break;
    default: cr = h * (c + s); sr = -h * (c - s); break;
  }
  *re = (double)cr;
  *im = (double)(-sr);  // exp(-i theta)
}

std::vector<double> stage_table(const SubPlan& sp, int s, bool transposed) {
  // power-of-two twiddles exp(-2 pi i np 2^k / len_s), k < NB, at [np*NB + k] (column pass) or
  // [k*stride + np] (row pass); fft_fixed.cuh:fx_twiddles builds the other powers from them
  const int R = sp.radix[s], st = sp.stride[s], len = sp.len[s];
  const int NB = R > 8 ? 4 : R > 4 ? 3 : R > 2 ? 2 : 1;
  std::vector<double> t(2 * (size_t)st * NB);
  for (int np = 0; np < st; ++np)
    for (int k = 0; k < NB; ++k) {
      double re, im;
      unit_root((int64_t)np * (1 << k), len, &re, &im);
      const size_t i = transposed ? (size_t)k * st + np : (size_t)np * NB + k;
      t[2 * i] = re;
      t[2 * i + 1] = im;
    }
  return t;
}

std::vector<double> subplan_table(const SubPlan& sp) {
  std::vector<double> t(2 * (size_t)sp.ntab);
  for (int i = 0; i < sp.ntab; ++i) {
    double re, im;
    unit_root(i, sp.S, &re, &im);
    if (sp.qmode) {  // (cos phi, sin phi)
      t[2 * i] = re;
      t[2 * i + 1] = -im;
    } else {  // exp(-2 pi i k / S)
      t[2 * i] = re;
      t[2 * i + 1] = im;
    }
  }
  return t;
}

static bool is_5smooth(int64_t n) {
  for (int p : {2, 3, 5})
    while (n % p == 0) n /= p;
  return n == 1;
}

static bool plan_2d_tile(const impb_ctx* ctx, int64_t n_min, bool exact, int elem_bytes, int tile, Plan2D* out);

bool plan_2d(const impb_ctx* ctx, int64_t n_min, bool exact, int elem_bytes, Plan2D* out, bool allow_half_tile) {
  // full tile = 128-byte HBM segments; transforms whose sub-FFTs do not fit shared memory at that width
  // (192 kHz sweeps) fall back to half-width tiles (64-byte segments)
  const int tile = 64 / elem_bytes;
  // A/B switch (tools/): plan the K1 transform on half-width tiles even where full-width ones fit
  static const bool force_half = getenv("IMPB_FORCE_HALF_TILE") != nullptr;
  if (force_half && allow_half_tile && plan_2d_tile(ctx, n_min, exact, elem_bytes, tile / 2, out)) return true;
  if (plan_2d_tile(ctx, n_min, exact, elem_bytes, tile, out)) {
    // a full-width plan that leaves room for ONE CTA per SM loses to half-width tiles when both sub-FFT lengths
    // have specialised kernels at that width (192 kHz: 1536 x 1600, two 256-thread CTAs of 98 KB per SM)
    const size_t smem_two = ((ctx && ctx->smem_per_sm) ? ctx->smem_per_sm : (size_t)228 * 1024) / 2 - 1024;
    Plan2D h;
    int fr[3];
    if (allow_half_tile && std::max(out->smem_col, out->smem_row) > smem_two &&
        plan_2d_tile(ctx, n_min, exact, elem_bytes, tile / 2, &h) && h.N1 >= 768 && h.N2 >= 768 &&
        fixed_radices(h.N1, fr) && fixed_radices(h.N2, fr) && fixed_tile_ok(h.N1, false) && fixed_tile_ok(h.N2, false))
      *out = h;
    return true;
  }
  return allow_half_tile && plan_2d_tile(ctx, n_min, exact, elem_bytes, tile / 2, out);
}

static bool plan_2d_tile(const impb_ctx* ctx, int64_t n_min, bool exact, int elem_bytes, int tile, Plan2D* out) {
  const size_t cx = 2 * (size_t)elem_bytes;
  const size_t smem_max = (ctx && ctx->smem_optin) ? ctx->smem_optin : (size_t)227 * 1024;
  const size_t smem_two = ((ctx && ctx->smem_per_sm) ? ctx->smem_per_sm : (size_t)228 * 1024) / 2 - 1024;
  if (n_min < 4) return false;

  int force1 = 0, force2 = 0;
  if (!exact) {
    const char* f = getenv("IMPB_PLAN");  // tuning hook: "N1xN2"
    if (f) sscanf(f, "%dx%d", &force1, &force2);
  }

  // candidate sub-FFT lengths
  struct Cand { int S; size_t smem; int threads; double cost; };
  std::vector<Cand> cands;
  for (int S = 2; S <= 4096; ++S) {
    if (!is_5smooth(S)) continue;
    const size_t ntab = (S % 4 == 0) ? S / 4 : S;
    const size_t smem = ((size_t)((S + 7) & ~7) * tile + ntab) * cx;
    if (smem > smem_max) break;
    const int threads = smem > smem_two ? 512 : 256;
    std::vector<int> rad;
    double c = subplan_cost(S, tile, threads, &rad);
    if (c >= 1e299) continue;
    if (smem > smem_two) c *= 1.3;  // one CTA per SM: less latency hiding
    int fr[3];
    if (fixed_radices(S, fr) && fixed_tile_ok(S, tile == 64 / elem_bytes)) c *= 0.4;  // compile-time specialised kernel
    cands.push_back({S, smem, threads, c});
  }
  double best = 1e300;
  int b1 = 0, b2 = 0;
  for (const Cand& c1 : cands) {
    for (const Cand& c2 : cands) {
      const int64_t N = (int64_t)c1.S * c2.S;
      if (force1 && (c1.S != force1 || c2.S != force2)) continue;
      if (exact ? (N != n_min) : (N < n_min || (!force1 && N > n_min + n_min / 2))) continue;
      // column pass: ceil(N2/tile) CTAs of plan c1; row pass: ceil(N1/tile) CTAs of plan c2
      const double ctas_col = (double)((c2.S + tile - 1) / tile);
      const double ctas_row = (double)((c1.S + tile - 1) / tile);
      double cost = ctas_col * c1.cost / c1.threads + ctas_row * c2.cost / c2.threads;
      if (!exact && (c1.S % 16 || c2.S % 16)) cost *= 1.05;  // unaligned rows
      if (cost < best) {
        best = cost;
        b1 = c1.S;
        b2 = c2.S;
      }
    }
  }
  if (!b1) return false;
  Plan2D p;
  p.N = (int64_t)b1 * b2;
  p.N1 = b1;
  p.N2 = b2;
  p.tile = tile;
  size_t s1 = 0, s2 = 0;
  int t1 = 256, t2 = 256;
  for (const Cand& c : cands) {
    if (c.S == b1) { s1 = c.smem; t1 = c.threads; }
    if (c.S == b2) { s2 = c.smem; t2 = c.threads; }
  }
  p.threads = std::max(t1, t2);
  if (!make_subplan(b1, tile, p.threads, &p.col) || !make_subplan(b2, tile, p.threads, &p.row)) return false;
  p.smem_col = s1;
  p.smem_row = s2;
  *out = p;
  return true;
}

}  // namespace impb

extern "C" {

uint32_t impb_version(void) { return 0x00010000u; }

const char* impb_last_error(void) { return impb::g_last_error.c_str(); }

int32_t impb_ctx_create(int32_t device, impb_ctx** out) {
  if (!out) return IMPB_EINVAL;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return impb::cuda_fail(e, "cudaGetDeviceCount (no CUDA device: this library has no CPU path)");
  if (device < 0 || device >= count) {
    impb::set_error("device %d out of range (%d devices)", device, count);
    return IMPB_EINVAL;
  }
  cudaDeviceProp prop;
  IMPB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    impb::set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                    prop.minor);
    return IMPB_EUNSUPPORTED;
  }
  // stream-ordered scratch (cudaMallocAsync) must stay cached across synchronisation points: with the default
  // release threshold of 0 every host sync returns the pool to the OS and the next launch pays a real allocation
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  impb_ctx* c = new impb_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  c->smem_per_sm = prop.sharedMemPerMultiprocessor;
  *out = c;
  return IMPB_OK;
}

void impb_ctx_destroy(impb_ctx* ctx) { delete ctx; }

int32_t impb_plan_probe(int64_t n_min, int32_t exact, int32_t elem_bytes, int64_t* n1, int64_t* n2,
                        int32_t* radices_col, int32_t* radices_row) {
  if ((elem_bytes != 8 && elem_bytes != 4) || n_min < 4) return IMPB_EINVAL;
  impb::Plan2D p;
  if (!impb::plan_2d(nullptr, n_min, exact != 0, elem_bytes, &p, true)) {
    impb::set_error("no plan for n=%lld", (long long)n_min);
    return IMPB_EUNSUPPORTED;
  }
  if (n1) *n1 = p.N1;
  if (n2) *n2 = p.N2;
  for (int s = 0; s < impb::MAX_STAGES; ++s) {
    if (radices_col) radices_col[s] = s < p.col.nst ? p.col.radix[s] : 0;
    if (radices_row) radices_row[s] = s < p.row.nst ? p.row.radix[s] : 0;
  }
  return IMPB_OK;
}

}  // extern "C"
