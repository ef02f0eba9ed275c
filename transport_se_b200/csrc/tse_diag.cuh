// Diagnostics: order-independent global tracer mass (the fixed-point idea of repro_sum, reference
// src/repro_sum_mod.F90:216-628, applied to the "Q mass" sum of prim_state_mod.F90:352-385).
// Each plane's J = sum_ij spheremp*Qdp is split into two int64 limbs relative to the global max exponent;
// integer adds commute, so the result is bitwise identical for any element order, block schedule or GPU count.
#pragma once
#include "tse_kernels.cuh"

namespace tse {

__device__ __forceinline__ double plane_mass(const Geo& G, const DssView& in, const ThreadPlane& t) {
  double v[16];
  in.load(G, t.e, t.q, t.k, v);
  const double* sp = G.spheremp + (size_t)t.e * 16;
  double J = 0.0;
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) J = fma(sp[n], v[n], J);
  return J;
}

// a warp holds 32 / SEG tracers of SEG planes each (SEG = 32 when a tracer's GPL planes fill whole warps)
constexpr int SEG = GPL < 32 ? GPL : 32;

__global__ void __launch_bounds__(GPL* QPB) k_mass_max(Geo G, DssView in, unsigned long long* __restrict__ maxbits) {
  const ThreadPlane t = thread_plane(G, in.Q);
  unsigned long long b = 0;
  if (t.valid) b = (unsigned long long)__double_as_longlong(fabs(plane_mass(G, in, t)));
  TSE_UNROLL
  for (int o = SEG / 2; o > 0; o >>= 1) {
    const unsigned long long x = __shfl_xor_sync(0xffffffffu, b, o);
    b = x > b ? x : b;
  }
  const int q = blockIdx.y * QPB + threadIdx.x / GPL;  // uniform per SEG lanes
  if ((threadIdx.x & (SEG - 1)) == 0 && q < in.Q) atomicMax(maxbits + q, b);
}

__global__ void __launch_bounds__(GPL* QPB) k_mass_fixed(Geo G, DssView in, const int* __restrict__ shift, long long* __restrict__ acc) {
  const ThreadPlane t = thread_plane(G, in.Q);
  const int q = blockIdx.y * QPB + threadIdx.x / GPL;
  long long hi = 0, lo = 0;
  if (t.valid) {
    const double x = scalbn(plane_mass(G, in, t), shift[q]);
    const double xi = trunc(x);
    hi = (long long)xi;
    lo = (long long)trunc(scalbn(x - xi, 40));
  }
  TSE_UNROLL
  for (int o = SEG / 2; o > 0; o >>= 1) {
    hi += __shfl_xor_sync(0xffffffffu, hi, o);
    lo += __shfl_xor_sync(0xffffffffu, lo, o);
  }
  if ((threadIdx.x & (SEG - 1)) == 0 && q < in.Q) {
    atomicAdd(reinterpret_cast<unsigned long long*>(acc + 2 * q), (unsigned long long)hi);
    atomicAdd(reinterpret_cast<unsigned long long*>(acc + 2 * q + 1), (unsigned long long)lo);
  }
}

// order-preserving map double -> uint64 (so that atomicMin/atomicMax on the integers order like the doubles)
__device__ __forceinline__ unsigned long long ordered_bits(double x) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// per-tracer global min/max of Q = Qdp / (dA*ps0 + dB*ps_v)  (the Q refresh of prim_driver_mod.F90:807-822 and the qmin/qmax lines
// of prim_printstate, prim_state_mod.F90:184-207)
__global__ void __launch_bounds__(GPL* QPB) k_q_minmax(Geo G, DssView in, const double* __restrict__ dA, const double* __restrict__ dB,
                                                       const double* __restrict__ ps_v, unsigned long long* __restrict__ omin,
                                                       unsigned long long* __restrict__ omax) {
  const ThreadPlane t = thread_plane(G, in.Q);
  double mn = 1e300, mx = -1e300;
  if (t.valid) {
    double v[16];
    in.load(G, t.e, t.q, t.k, v);
    const double* ps = ps_v + (size_t)t.e * 16;
    const double a = dA[t.k], b = dB[t.k];
    TSE_UNROLL
    for (int n = 0; n < 16; ++n) {
      const double q = v[n] / (a + b * ps[n]);
      mn = dmin(mn, q);
      mx = dmax(mx, q);
    }
  }
  TSE_UNROLL
  for (int o = SEG / 2; o > 0; o >>= 1) {
    mn = dmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = dmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int q = blockIdx.y * QPB + threadIdx.x / GPL;
  if ((threadIdx.x & (SEG - 1)) == 0 && q < in.Q) {
    atomicMin(omin + q, ordered_bits(mn));
    atomicMax(omax + q, ordered_bits(mx));
  }
}

}  // namespace tse
