// Per-thread (one 4x4 plane in registers) element operators of the tracer path.
//
// One thread owns all 16 GLL nodes of a (element, level, tracer) plane, so the 4x4
// derivative-matrix contractions, the limiter sums and the min/max reductions are
// register-only: no shuffles, no shared-memory traffic in the arithmetic.
//
// Reference semantics (citations relative to the reference tree):
//   divergence_sphere      src/share/derivative_mod.F90:2364-2414
//   gradient_sphere        src/share/derivative_mod.F90:1660-1700
//   divergence_sphere_wk   src/share/derivative_mod.F90:2027-2097
//   laplace_sphere_wk      src/share/derivative_mod.F90:2418-2460 (constant-coefficient branch)
//   limiter_optim_iter_full src/share/prim_advection_mod.F90:976-1094
// Node index n = i + 4*j (Fortran (i,j), i fastest).  Dvv(i,l) is D.d[i + 4*l].
//
// Algebraic folding used here (differs from the reference expression tree by O(1 ulp),
// far inside the 1e-12 parity bar, identical on every GPU count):
//   gv_c = metdet*(Dinv(c,1)*Vstar1*Qdp + Dinv(c,2)*Vstar2*Qdp) = U_c * Qdp,
//          U_c = metdet*(Dinv(c,1)*Vstar1 + Dinv(c,2)*Vstar2)           (per element, level; tracer independent)
//   laplace_sphere_wk: vtemp = Dinv Dinv^T (v1,v2), so with T = spheremp*rrearth^2*Dinv*Dinv^T (symmetric, 3 numbers
//          per node)  lap(m,n) = -sum_j [ w1(j,n) Dvv(m,j) + w2(m,j) Dvv(n,j) ],  w = T (d1,d2)
#pragma once
#include "tse_layout.cuh"

namespace tse {

struct Dvv {
  double d[16];
};

#define TSE_UNROLL _Pragma("unroll")

// min/max as compare+select (DSETP + SEL).  fmin/fmax expand to ~8 instructions on sm_100a (NaN-quieting sequence);
// NaNs do not occur on this path, and for ordinary numbers the results are identical.
__device__ __forceinline__ double dmin(double a, double b) { return b < a ? b : a; }
__device__ __forceinline__ double dmax(double a, double b) { return b > a ? b : a; }

// r(a,b) = sum_i Dvv(i,a) g1(i,b) + sum_i Dvv(i,b) g2(a,i)      [div(l,j) + vvtemp(j,l) of divergence_sphere]
__device__ __forceinline__ void div_contract(const double (&g1)[16], const double (&g2)[16], const Dvv& D, double (&r)[16]) {
  TSE_UNROLL
  for (int b = 0; b < 4; ++b) {
    TSE_UNROLL
    for (int a = 0; a < 4; ++a) {
      double s1 = 0.0, s2 = 0.0;
      TSE_UNROLL
      for (int i = 0; i < 4; ++i) {
        // Dvv(i,i) == 0 for the two interior nodes (derivative_mod.F90:451-486): adding 0*g is exact, skip it
        if (!(i == a && (a == 1 || a == 2))) s1 = fma(D.d[i + 4 * a], g1[i + 4 * b], s1);
        if (!(i == b && (b == 1 || b == 2))) s2 = fma(D.d[i + 4 * b], g2[a + 4 * i], s2);
      }
      r[a + 4 * b] = s1 + s2;
    }
  }
}

// raw derivative sums of gradient_sphere: d1(a,b) = sum_i Dvv(i,a) s(i,b), d2(a,b) = sum_i Dvv(i,b) s(a,i)
__device__ __forceinline__ void grad_raw(const double (&s)[16], const Dvv& D, double (&d1)[16], double (&d2)[16]) {
  TSE_UNROLL
  for (int b = 0; b < 4; ++b) {
    TSE_UNROLL
    for (int a = 0; a < 4; ++a) {
      double s1 = 0.0, s2 = 0.0;
      TSE_UNROLL
      for (int i = 0; i < 4; ++i) {
        if (!(i == a && (a == 1 || a == 2))) s1 = fma(D.d[i + 4 * a], s[i + 4 * b], s1);
        if (!(i == b && (b == 1 || b == 2))) s2 = fma(D.d[i + 4 * b], s[a + 4 * i], s2);
      }
      d1[a + 4 * b] = s1;
      d2[a + 4 * b] = s2;
    }
  }
}

// laplace_sphere_wk of s with the per-node tensor T (t11,t12,t22 point at this element's 16 nodes)
__device__ __forceinline__ void laplace_wk(const double (&s)[16], const Dvv& D, const double* __restrict__ t11,
                                           const double* __restrict__ t12, const double* __restrict__ t22, double (&lap)[16]) {
  double w1[16], w2[16];
  {
    double d1[16], d2[16];
    grad_raw(s, D, d1, d2);
    TSE_UNROLL
    for (int n = 0; n < 16; ++n) {
      const double a = t11[n], b = t12[n], c = t22[n];
      w1[n] = a * d1[n] + b * d2[n];
      w2[n] = b * d1[n] + c * d2[n];
    }
  }
  TSE_UNROLL
  for (int nn = 0; nn < 4; ++nn) {
    TSE_UNROLL
    for (int m = 0; m < 4; ++m) {
      double acc = 0.0;
      TSE_UNROLL
      for (int j = 0; j < 4; ++j) {
        if (!(j == m && (m == 1 || m == 2))) acc = fma(-w1[j + 4 * nn], D.d[m + 4 * j], acc);
        if (!(j == nn && (nn == 1 || nn == 2))) acc = fma(-w2[m + 4 * j], D.d[nn + 4 * j], acc);
      }
      lap[m + 4 * nn] = acc;
    }
  }
}

// limiter_optim_iter_full.  On entry x = ptens (tracer mass), c = sphweights*dpmass, rdpm = 1/dpmass.
// On exit x = limited mixing ratio (the reference's x before "ptens = x*dpmass"); minp/maxp are updated
// in place exactly as the reference relaxes them (prim_advection_mod.F90:1024-1029).
// The sums run in the reference's k1 order (i outer, j inner, :1006-1013).
__device__ __forceinline__ void limiter_optim_iter_full(double (&x)[16], const double (&c)[16], const double (&rdpm)[16],
                                                        double& minp, double& maxp) {
  const double tol_limiter = (double)5e-14f;  // default-real literal in the reference (:1003)
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) x[n] = x[n] * rdpm[n];
  double sumc = 0.0, mass = 0.0;
  TSE_UNROLL
  for (int k1 = 0; k1 < 16; ++k1) {
    const int n = (k1 >> 2) + 4 * (k1 & 3);
    sumc += c[n];
  }
  if (sumc <= 0.0) return;  // :1016 (x*c == ptens*sphweights up to rounding)
  TSE_UNROLL
  for (int k1 = 0; k1 < 16; ++k1) {
    const int n = (k1 >> 2) + 4 * (k1 & 3);
    mass = fma(c[n], x[n], mass);
  }
  if (mass < minp * sumc) minp = mass / sumc;
  if (mass > maxp * sumc) maxp = mass / sumc;

  const double thresh = tol_limiter * fabs(mass);
  for (int iter = 1; iter <= NPSQ - 1; ++iter) {
    double addmass = 0.0;
    TSE_UNROLL
    for (int k1 = 0; k1 < 16; ++k1) {
      const int n = (k1 >> 2) + 4 * (k1 & 3);
      if (x[n] > maxp) {
        addmass = fma(x[n] - maxp, c[n], addmass);
        x[n] = maxp;
      }
      if (x[n] < minp) {
        addmass = fma(-(minp - x[n]), c[n], addmass);
        x[n] = minp;
      }
    }
    if (fabs(addmass) <= thresh) break;
    double weightssum = 0.0;
    if (addmass > 0.0) {
      TSE_UNROLL
      for (int k1 = 0; k1 < 16; ++k1) {
        const int n = (k1 >> 2) + 4 * (k1 & 3);
        if (x[n] < maxp) weightssum += c[n];
      }
      const double inc = addmass / weightssum;
      TSE_UNROLL
      for (int n = 0; n < 16; ++n)
        if (x[n] < maxp) x[n] += inc;
    } else {
      TSE_UNROLL
      for (int k1 = 0; k1 < 16; ++k1) {
        const int n = (k1 >> 2) + 4 * (k1 & 3);
        if (x[n] > minp) weightssum += c[n];
      }
      const double inc = addmass / weightssum;
      TSE_UNROLL
      for (int n = 0; n < 16; ++n)
        if (x[n] > minp) x[n] += inc;
    }
  }
}

__device__ __forceinline__ void load16(const double* __restrict__ p, double (&v)[16]) {
  const double2* s = reinterpret_cast<const double2*>(p);
  TSE_UNROLL
  for (int c = 0; c < 8; ++c) {
    const double2 t = s[c];
    v[2 * c] = t.x;
    v[2 * c + 1] = t.y;
  }
}
__device__ __forceinline__ void store16(double* __restrict__ p, const double (&v)[16]) {
  double2* s = reinterpret_cast<double2*>(p);
  TSE_UNROLL
  for (int c = 0; c < 8; ++c) s[c] = make_double2(v[2 * c], v[2 * c + 1]);
}

}  // namespace tse
