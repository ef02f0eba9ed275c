// Diagnostics on the plane-per-thread view: global extrema of Q and the field fingerprint.  (The global tracer mass runs through
// the tile pipeline, OP_MASS in tse_pipe.cuh.)
#pragma once
#include "tse_kernels.cuh"

namespace tse {

// a warp holds 32 / SEG tracers of SEG planes each (SEG = 32 when a tracer's GPL planes fill whole warps)
constexpr int SEG = GPL < 32 ? GPL : 32;

// Field fingerprint for bit-for-bit comparisons across GPU counts: per tracer, the wrapping 64-bit sum over all (element, level,
// node) of mix(bit pattern of Qdp, global position).  The position key uses the element's global space-filling-curve index, so
// the value does not depend on which rank owns the element or where it sits in memory; integer sums commute, so it does not
// depend on the order either.  Two runs give equal fingerprints iff (up to 2^-64 collisions) every value of the field is bitwise
// equal at every global position.
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {  // splitmix64 finaliser
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}
__global__ void __launch_bounds__(GPL* QPB) k_field_hash(Geo G, DssView in, const int* __restrict__ gkey, unsigned long long* __restrict__ acc) {
  const ThreadPlane t = thread_plane(G, in.Q);
  unsigned long long h = 0;
  if (t.valid) {
    double v[16];
    in.load(G, t.e, t.q, t.k, v);
    const unsigned long long pos = (((unsigned long long)gkey[t.e] * NLEV + t.k) * 4096ull + t.q) * 16ull;
    TSE_UNROLL
    for (int n = 0; n < 16; ++n) h += mix64((unsigned long long)__double_as_longlong(v[n]) ^ mix64(pos + n));
  }
  TSE_UNROLL
  for (int o = SEG / 2; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
  const int q = blockIdx.y * QPB + threadIdx.x / GPL;  // uniform per SEG lanes
  if ((threadIdx.x & (SEG - 1)) == 0 && q < in.Q) atomicAdd(acc + q, h);
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
