// Shared host-side plumbing of libimpulcifer_b200.so: context, error reporting, FFT planning.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/impb.h"
#include "fft_tile.cuh"

struct impb_ctx {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;   // max dynamic shared memory per CTA
  size_t smem_per_sm = 0;
};

namespace impb {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define IMPB_CUDA(call)                                          \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) return ::impb::cuda_fail(e__, #call); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur;
    cudaGetDevice(&cur);
    if (cur != prev && prev >= 0) cudaSetDevice(prev);
  }
};

// ---- planning ------------------------------------------------------------------------------------
// Factor S into in-register radices; returns false if S has a prime factor > 5 or does not fit.
bool make_subplan(int S, int tile, int threads, SubPlan* sp);
// radices of the compile-time specialised kernel for length S, if one exists
bool fixed_radices(int S, int* r);
// ... and whether those kernels exist for full-width (true) or half-width (false) tiles
bool fixed_tile_ok(int S, bool full_tile);
// per-stage twiddle table of the specialised kernels (column-pass or transposed row-pass layout)
std::vector<double> stage_table(const SubPlan& sp, int s, bool transposed);
// the same with every power p = 1..R-1 stored (exact twiddles for the frequency-response transforms)
std::vector<double> stage_table_full(const SubPlan& sp, int s, bool transposed);
// digit-reversal: perm[pos] = frequency index held at tile position pos after the DIF stages
std::vector<int> subplan_perm(const SubPlan& sp);

struct Plan2D {
  int64_t N = 0;
  int N1 = 0, N2 = 0;       // N = N1 * N2; column pass length N1 (stride N2), row pass length N2
  SubPlan col, row;
  int tile = 0;             // columns / rows per CTA
  int threads = 0;
  size_t smem_col = 0, smem_row = 0;
};
// smallest-cost N = N1*N2 >= n_min (exact = true: N must equal n_min) for element type of `elem_bytes`
// (8 = double, 4 = float); false if impossible.
bool plan_2d(const impb_ctx* ctx, int64_t n_min, bool exact, int elem_bytes, Plan2D* out, bool allow_half_tile = false);

// quarter-wave (or full) twiddle table of a sub-plan in double precision: (cos, sin) pairs
std::vector<double> subplan_table(const SubPlan& sp);
// exp(-2 pi i k / N) evaluated in extended precision
void unit_root(int64_t k, int64_t N, double* re, double* im);

}  // namespace impb
