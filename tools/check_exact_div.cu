// Developer check (not shipped): ExactDivisor(d)(s) == __ddiv_rn(s, d) bit for bit on 2^33 pseudo-random pairs,
// including mantissas of all ones / one-above-a-power-of-two, |s| <= |d| and |s| > |d|, and the band edges.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/check_exact_div tools/check_exact_div.cu && build/check_exact_div
#include <cstdint>
#include <cstdio>

#include "../impulcifer_pip313_b200/csrc/exact_div.cuh"

__device__ __forceinline__ uint64_t mix(uint64_t z) {
  z += 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

// a double with a random sign, an exponent drawn from [emin, emax] and a mantissa from one of four families
__device__ double draw(uint64_t h, int emin, int emax) {
  const uint64_t sign = (h >> 63) << 63;
  const int e = emin + (int)((h >> 40) % (uint64_t)(emax - emin + 1));
  uint64_t m = mix(h) & 0xfffffffffffffull;
  switch ((h >> 20) & 7) {
    case 0: m = 0xfffffffffffffull - (m & 7); break;   // all ones (minus a few ulps)
    case 1: m = m & 7; break;                          // just above a power of two
    case 2: m &= 0xfffff00000000ull; break;            // short mantissa
    default: break;
  }
  return __longlong_as_double((long long)(sign | ((uint64_t)(e + 1023) << 52) | m));
}

__global__ void k_check(uint64_t seed, int rows, int per_row, unsigned long long* bad, double* first_bad) {
  for (int r = blockIdx.x; r < rows; r += gridDim.x) {
    const uint64_t hr = mix(seed * 0x10001ull + (uint64_t)r);
    const int band = (int)(hr & 3);
    // band 0: audio-like (|d| around 1e-3..1e3), 1: wide, 2: edges of the fast band, 3: outside it
    const double d = band == 0 ? draw(hr, -12, 12) : band == 1 ? draw(hr, -300, 300) : band == 2 ? draw(hr, 325, 335)
                                                                                                 : draw(hr, -1000, -330);
    const impb::ExactDivisor div(d);
    for (int i = threadIdx.x; i < per_row; i += blockDim.x) {
      const uint64_t h = mix(hr ^ ((uint64_t)i * 0x2545f4914f6cdd1dull));
      double s;
      switch (h & 7) {
        case 0: s = draw(h, -60, 0) * d; break;                       // |s| <= |d| (the Lundeby case), rounded product
        case 1: s = draw(h, -520, -480); break;                       // around the 1e-150 edge
        case 2: s = draw(h, 480, 520); break;                         // around the 1e150 edge
        case 3: s = draw(h, -1022, 1023); break;                      // anywhere
        case 4: s = (h >> 8) & 1 ? 0.0 : -0.0; break;
        default: s = draw(h, -40, 40); break;
      }
      const double got = div(s), want = __ddiv_rn(s, d);
      if (__double_as_longlong(got) != __double_as_longlong(want) && !(got != got && want != want)) {
        if (atomicAdd(bad, 1ull) == 0) { first_bad[0] = s; first_bad[1] = d; first_bad[2] = got; first_bad[3] = want; }
      }
    }
  }
}

int main() {
  unsigned long long* bad;
  double* fb;
  cudaMallocManaged(&bad, 8);
  cudaMallocManaged(&fb, 32);
  *bad = 0;
  const int rows = 1 << 15, per_row = 1 << 14;
  unsigned long long total = 0;
  for (uint64_t seed = 1; seed <= 16; ++seed) {
    k_check<<<148 * 8, 256>>>(seed, rows, per_row, bad, fb);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 2; }
    total += (unsigned long long)rows * per_row;
  }
  printf("pairs %llu mismatches %llu\n", total, *bad);
  if (*bad) printf("first: s=%a d=%a got=%a want=%a\n", fb[0], fb[1], fb[2], fb[3]);
  return *bad ? 1 : 0;
}
