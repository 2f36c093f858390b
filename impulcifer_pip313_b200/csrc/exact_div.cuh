// Correctly rounded s / d for MANY numerators over ONE denominator: the reciprocal is rounded once per row and every
// quotient costs five FP64 operations instead of a full division.
//   y = RN(1/d);  q0 = RN(s y);  r0 = s - d q0 (exact, FMA);  q1 = RN(q0 + r0 y);  r1 = s - d q1;  q = RN(q1 + r1 y)
// q1 is within one ulp of s/d, so the second correction rounds correctly (Markstein, "Computation of elementary
// functions on the IBM RISC System/6000 processor", 1990, theorem 2: y the correctly rounded reciprocal, q1 faithful,
// r1 exact).  The residuals are only exact away from underflow and overflow, so the fast form is used when d and s are
// inside a wide safe band and __ddiv_rn otherwise; tools/check_exact_div.cu compares the two on the GPU.
#pragma once

namespace impb {

struct ExactDivisor {
  double d, y;
  bool fast;
  __device__ __forceinline__ explicit ExactDivisor(double d_) : d(d_) {
    const double a = fabs(d_);
    fast = a >= 1e-100 && a <= 1e100;
    y = fast ? __drcp_rn(d_) : 0.0;
  }
  __device__ __forceinline__ double operator()(double s) const {
    const double a = fabs(s);
    if (fast && a >= 1e-150 && a <= 1e150) {
      const double q0 = __dmul_rn(s, y);
      const double q1 = __fma_rn(__fma_rn(-d, q0, s), y, q0);
      return __fma_rn(__fma_rn(-d, q1, s), y, q1);
    }
    return __ddiv_rn(s, d);
  }
};

}  // namespace impb
