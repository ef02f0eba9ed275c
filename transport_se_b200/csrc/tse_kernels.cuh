// CUDA kernels of the tracer-advection path (sm_100a, FP64).
//
// Work decomposition: one thread owns one 4x4 plane (element, level, tracer); a CTA covers
// (group of 16 elements, chunk of 4 levels) x QPB tracers, i.e. QPB contiguous 8 KB tiles of the
// tracer field.  The DSS (edgeVpack / bndry_exchangeV / edgeVunpack, edge_mod.F90:366-742) is never
// materialised as an edge buffer: a kernel that consumes a field produced "pre-DSS" gathers the
// neighbour nodes in the reference's unpack order while loading (see DssView::resolve).
#pragma once
#include "tse_ops.cuh"

namespace tse {

constexpr int QPB = 4;  // tracers per CTA in the plane-per-thread kernels

struct Geo {
  const double* spheremp;   // [e][16]
  const double* rspheremp;  // [e][16]
  const double* rmp;        // 1/spheremp
  const double* rmr;        // rmetdet*rrearth
  const double* mD;         // [e][4][16]: metdet*Dinv(1,1), (1,2), (2,1), (2,2)
  const double* T;          // [e][3][16]: spheremp*rrearth^2*(Dinv Dinv^T) 11,12,22
  const int* gsrc;          // [e][NSLOT] DSS gather table
  const int* nbr8;          // [e][8] neighbour elements in unpack order S,E,N,W,SW,SE,NE,NW (>=0 local, -1 none, <=-2 ghost bundle)
  int nelem;                // real elements; arrays are padded to ngroups*GE
  int ngroups;
};

// Read-side view of a tracer field that may still need its DSS applied.
struct DssView {
  const double* q;      // field [g][kc][q][el][kk][16]
  const double* ghost;  // halo values of off-GPU neighbours: [slot][q][k]
  int pending;          // 1: q holds pre-DSS values (spheremp-weighted), gather + rspheremp on load
  int Q;                // tracers in this field

  // v = rspheremp * (own + neighbours) in the reference's order: S,E,N,W edges then SW,SE,NE,NW corners
  __device__ __forceinline__ void load(const Geo& G, int e, int q_, int k, double (&v)[16]) const {
    load16(q + qplane(e, q_, k, Q) * 16, v);
    if (!pending) return;
    const int* gs = G.gsrc + (size_t)e * NSLOT;
    const int kc = k / KC, kk = k % KC;
    auto fetch = [&](int s) -> double {
      if (s >= 0) {
        const int es = s >> 4, g = es / GE, el = es % GE;
        return q[(((((size_t)g * NKC + kc) * Q + q_) * GE + el) * KC + kk) * 16 + (s & 15)];
      }
      return ghost[((size_t)(-s - 2) * Q + q_) * NLEV + k];
    };
    TSE_UNROLL
    for (int t = 0; t < 4; ++t) { const int s = gs[t]; if (s != -1) v[t] += fetch(s); }
    TSE_UNROLL
    for (int t = 0; t < 4; ++t) { const int s = gs[4 + t]; if (s != -1) v[3 + 4 * t] += fetch(s); }
    TSE_UNROLL
    for (int t = 0; t < 4; ++t) { const int s = gs[8 + t]; if (s != -1) v[12 + t] += fetch(s); }
    TSE_UNROLL
    for (int t = 0; t < 4; ++t) { const int s = gs[12 + t]; if (s != -1) v[4 * t] += fetch(s); }
    { const int s = gs[16]; if (s != -1) v[0] += fetch(s); }
    { const int s = gs[17]; if (s != -1) v[3] += fetch(s); }
    { const int s = gs[18]; if (s != -1) v[15] += fetch(s); }
    { const int s = gs[19]; if (s != -1) v[12] += fetch(s); }
    const double* rs = G.rspheremp + (size_t)e * 16;
    TSE_UNROLL
    for (int n = 0; n < 16; ++n) v[n] = rs[n] * v[n];
  }
};

// Per-(element, level) stage package written by k_stage_prep: 5 planes per level plane.
//   0: U1   1: U2   2: rdp = 1/dp_s   3: c = spheremp*dp_star   4: rdpstar = 1/dp_star
// dp_s = derived%dp - rhs_multiplier*dt*divdp_proj (prim_advection_mod.F90:753,847), dp_star = dp_s - dt*divdp (:864)
constexpr int NPKG = 5;

struct ThreadPlane {
  int e, q, k;
  bool valid;
};
__device__ __forceinline__ ThreadPlane thread_plane(const Geo& G, int Q) {
  ThreadPlane t;
  const int gk = blockIdx.x, g = gk / NKC, kc = gk % NKC;
  const int tid = threadIdx.x;
  t.q = blockIdx.y * QPB + tid / GPL;
  const int el = (tid / KC) % GE, kk = tid % KC;
  t.e = g * GE + el;
  t.k = kc * KC + kk;
  t.valid = (t.e < G.nelem) && (t.q < Q);
  return t;
}

// ---------------------------------------------------------------------------------------------
// level-field kernels (one thread per (element, level) plane)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool thread_level(const Geo& G, int& e, int& k) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // level plane index in layout order
  const int kk = idx % KC, el = (idx / KC) % GE, gk = idx / GPL;
  const int g = gk / NKC, kc = gk % NKC;
  e = g * GE + el;
  k = kc * KC + kk;
  return g < G.ngroups && e < G.nelem;
}

// divdp = divergence_sphere(vn0); divdp_proj = divdp   (prim_advection_mod.F90:614-623)
__global__ void __launch_bounds__(128) k_divdp(Geo G, Dvv D, const double* __restrict__ vn0, double* __restrict__ divdp,
                                               double* __restrict__ divdp_proj) {
  int e, k;
  if (!thread_level(G, e, k)) return;
  double v1[16], v2[16], g1[16], g2[16], r[16];
  load16(vn0 + vplane(e, k, 0) * 16, v1);
  load16(vn0 + vplane(e, k, 1) * 16, v2);
  const double* mD = G.mD + (size_t)e * 64;
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) {
    g1[n] = mD[n] * v1[n] + mD[16 + n] * v2[n];
    g2[n] = mD[32 + n] * v1[n] + mD[48 + n] * v2[n];
  }
  div_contract(g1, g2, D, r);
  const double* rmr = G.rmr + (size_t)e * 16;
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) r[n] = r[n] * rmr[n];
  const size_t lp = lplane(e, k) * 16;
  store16(divdp + lp, r);
  store16(divdp_proj + lp, r);
}

// stage package (see NPKG)
__global__ void __launch_bounds__(128) k_stage_prep(Geo G, const double* __restrict__ vn0, const double* __restrict__ dp,
                                                    const double* __restrict__ divdp, const double* __restrict__ divdp_proj,
                                                    double rhs_mult_dt, double dt, double* __restrict__ pkg, size_t pkg_stride) {
  int e, k;
  if (!thread_level(G, e, k)) return;
  const size_t lp = lplane(e, k) * 16;
  double v1[16], v2[16], d[16], dd[16], dj[16];
  load16(vn0 + vplane(e, k, 0) * 16, v1);
  load16(vn0 + vplane(e, k, 1) * 16, v2);
  load16(dp + lp, d);
  load16(divdp + lp, dd);
  load16(divdp_proj + lp, dj);
  const double* mD = G.mD + (size_t)e * 64;
  const double* sp = G.spheremp + (size_t)e * 16;
  double o[16];
  double rdp[16], dps[16];
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) {
    const double dp_s = d[n] - rhs_mult_dt * dj[n];
    rdp[n] = 1.0 / dp_s;
    dps[n] = dp_s - dt * dd[n];
    v1[n] = v1[n] * rdp[n];  // Vstar
    v2[n] = v2[n] * rdp[n];
  }
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) o[n] = mD[n] * v1[n] + mD[16 + n] * v2[n];
  store16(pkg + 0 * pkg_stride + lp, o);
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) o[n] = mD[32 + n] * v1[n] + mD[48 + n] * v2[n];
  store16(pkg + 1 * pkg_stride + lp, o);
  store16(pkg + 2 * pkg_stride + lp, rdp);
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) o[n] = sp[n] * dps[n];
  store16(pkg + 3 * pkg_stride + lp, o);
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) o[n] = 1.0 / dps[n];
  store16(pkg + 4 * pkg_stride + lp, o);
}

// DSS of one level field (the DSSopt variable of euler_step, prim_advection_mod.F90:913-919,943-958):
// out = rspheremp * sum_{sharing elements} spheremp*f, in unpack order.  ghost: [slot][k] = spheremp*f of off-GPU nodes.
__global__ void __launch_bounds__(128) k_dss_level(Geo G, const double* __restrict__ f, const double* __restrict__ ghost,
                                                   double* __restrict__ out) {
  int e, k;
  if (!thread_level(G, e, k)) return;
  double v[16];
  load16(f + lplane(e, k) * 16, v);
  const double* sp = G.spheremp + (size_t)e * 16;
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) v[n] = sp[n] * v[n];
  const int* gs = G.gsrc + (size_t)e * NSLOT;
  auto fetch = [&](int s) -> double {
    if (s >= 0) {
      const int es = s >> 4, nd = s & 15;
      return G.spheremp[(size_t)es * 16 + nd] * f[lplane(es, k) * 16 + nd];
    }
    return ghost[(size_t)(-s - 2) * NLEV + k];
  };
  TSE_UNROLL
  for (int t = 0; t < 4; ++t) { const int s = gs[t]; if (s != -1) v[t] += fetch(s); }
  TSE_UNROLL
  for (int t = 0; t < 4; ++t) { const int s = gs[4 + t]; if (s != -1) v[3 + 4 * t] += fetch(s); }
  TSE_UNROLL
  for (int t = 0; t < 4; ++t) { const int s = gs[8 + t]; if (s != -1) v[12 + t] += fetch(s); }
  TSE_UNROLL
  for (int t = 0; t < 4; ++t) { const int s = gs[12 + t]; if (s != -1) v[4 * t] += fetch(s); }
  { const int s = gs[16]; if (s != -1) v[0] += fetch(s); }
  { const int s = gs[17]; if (s != -1) v[3] += fetch(s); }
  { const int s = gs[18]; if (s != -1) v[15] += fetch(s); }
  { const int s = gs[19]; if (s != -1) v[12] += fetch(s); }
  const double* rs = G.rspheremp + (size_t)e * 16;
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) v[n] = v[n] * rs[n];
  store16(out + lplane(e, k) * 16, v);
}

// ---------------------------------------------------------------------------------------------
// tracer-field kernels (one thread per (element, level, tracer) plane)
// ---------------------------------------------------------------------------------------------
struct MinMaxIO {
  double* qmin;            // [plane] limiter bounds, in/out (prim_advection_mod.F90:461 qmin/qmax(nlev,qsize,nelemd))
  double* qmax;
  double* qmin_loc;        // [plane] element-local extrema written before a neighbour exchange
  double* qmax_loc;
  const double* ghost_mm;  // [bundle][2][q][k] extrema of off-GPU neighbour elements
};

// element-local min/max of Q = Qdp/dp (prim_advection_mod.F90:764-775): feeds neighbor_minmax
__global__ void __launch_bounds__(GPL* QPB) k_minmax_local(Geo G, DssView in, const double* __restrict__ pkg, size_t pkg_stride,
                                                           MinMaxIO mm) {
  const ThreadPlane t = thread_plane(G, in.Q);
  if (!t.valid) return;
  double v[16], rdp[16];
  in.load(G, t.e, t.q, t.k, v);
  load16(pkg + 2 * pkg_stride + lplane(t.e, t.k) * 16, rdp);
  double mn = v[0] * rdp[0], mx = mn;
  TSE_UNROLL
  for (int n = 1; n < 16; ++n) {
    const double qv = v[n] * rdp[n];
    mn = fmin(mn, qv);
    mx = fmax(mx, qv);
  }
  const size_t p = qplane(t.e, t.q, t.k, in.Q);
  mm.qmin_loc[p] = mn;
  mm.qmax_loc[p] = mx;
}

// min/max over the element and its (up to 8) neighbours: neighbor_minmax, viscosity_mod.F90:748-816
__device__ __forceinline__ void neighbor_minmax(const Geo& G, const MinMaxIO& mm, int e, int q, int k, int Q, double& mn, double& mx) {
  const size_t p = qplane(e, q, k, Q);
  mn = mm.qmin_loc[p];
  mx = mm.qmax_loc[p];
  const int* nb = G.nbr8 + (size_t)e * 8;
  TSE_UNROLL
  for (int d = 0; d < 8; ++d) {
    const int b = nb[d];
    if (b >= 0) {
      const size_t pb = qplane(b, q, k, Q);
      mn = fmin(mn, mm.qmin_loc[pb]);
      mx = fmax(mx, mm.qmax_loc[pb]);
    } else if (b <= -2) {
      const size_t gb = ((size_t)(-b - 2) * 2 * Q + q) * NLEV + k;
      mn = fmin(mn, mm.ghost_mm[gb]);
      mx = fmax(mx, mm.ghost_mm[gb + (size_t)Q * NLEV]);
    }
  }
}

// first half of biharmonic_wk_scalar_minmax (viscosity_mod.F90:353-405): Q = Qdp/dp, local extrema, qtens = laplace_sphere_wk(Q)
__global__ void __launch_bounds__(GPL* QPB) k_biharm_pre(Geo G, Dvv D, DssView in, const double* __restrict__ pkg, size_t pkg_stride,
                                                         MinMaxIO mm, double* __restrict__ qtens) {
  const ThreadPlane t = thread_plane(G, in.Q);
  if (!t.valid) return;
  double v[16], lap[16];
  in.load(G, t.e, t.q, t.k, v);
  {
    double rdp[16];
    load16(pkg + 2 * pkg_stride + lplane(t.e, t.k) * 16, rdp);
    TSE_UNROLL
    for (int n = 0; n < 16; ++n) v[n] = v[n] * rdp[n];
  }
  double mn = v[0], mx = v[0];
  TSE_UNROLL
  for (int n = 1; n < 16; ++n) {
    mn = fmin(mn, v[n]);
    mx = fmax(mx, v[n]);
  }
  const size_t p = qplane(t.e, t.q, t.k, in.Q);
  mm.qmin_loc[p] = mn;
  mm.qmax_loc[p] = mx;
  const double* T = G.T + (size_t)t.e * 48;
  laplace_wk(v, D, T, T + 16, T + 32, lap);
  store16(qtens + p * 16, lap);
}

// One RK stage of euler_step (prim_advection_mod.F90:667-970) for rhs_multiplier = MODE-1.
//   MODE 1: bounds = neighbor_minmax of the pre-computed local extrema
//   MODE 2: bounds = min/max(stored bounds, local extrema of the DSS'd input)        (:781-793)
//   MODE 3: bounds = neighbour extrema; adds Qtens_biharmonic = -3*dt*nu_q*dp0(k)*lap(rspheremp*DSS(qtens))/spheremp (:796-827)
// Output: Qdp(np1) = spheremp*limited(Qtens), still to be DSS'd (the consumer gathers).
struct StageArgs {
  DssView in;
  DssView qtens;  // MODE 3 only
  const double* pkg;
  size_t pkg_stride;
  MinMaxIO mm;
  double* out;
  double dt;
  double visc_coef;  // -rhs_viss*dt*nu_q
  const double* dp0; // [NLEV] (hyai(k+1)-hyai(k))*ps0 + (hybi(k+1)-hybi(k))*ps0
};

template <int MODE>
__global__ void __launch_bounds__(GPL* QPB) k_euler_stage(Geo G, Dvv D, StageArgs a) {
  const int Q = a.in.Q;
  const ThreadPlane t = thread_plane(G, Q);
  if (!t.valid) return;
  const size_t p = qplane(t.e, t.q, t.k, Q);
  const size_t lp = lplane(t.e, t.k) * 16;
  double v[16];
  a.in.load(G, t.e, t.q, t.k, v);

  double minp, maxp;
  if (MODE == 2) {
    double rdp[16];
    load16(a.pkg + 2 * a.pkg_stride + lp, rdp);
    double mn = v[0] * rdp[0], mx = mn;
    TSE_UNROLL
    for (int n = 1; n < 16; ++n) {
      const double qv = v[n] * rdp[n];
      mn = fmin(mn, qv);
      mx = fmax(mx, qv);
    }
    minp = fmin(a.mm.qmin[p], mn);
    maxp = fmax(a.mm.qmax[p], mx);
  } else {
    neighbor_minmax(G, a.mm, t.e, t.q, t.k, Q, minp, maxp);
  }

  double x[16];
  {
    double g1[16], g2[16];
    {
      double u[16];
      load16(a.pkg + lp, u);
      TSE_UNROLL
      for (int n = 0; n < 16; ++n) g1[n] = u[n] * v[n];
      load16(a.pkg + a.pkg_stride + lp, u);
      TSE_UNROLL
      for (int n = 0; n < 16; ++n) g2[n] = u[n] * v[n];
    }
    div_contract(g1, g2, D, x);
    const double* rmr = G.rmr + (size_t)t.e * 16;
    TSE_UNROLL
    for (int n = 0; n < 16; ++n) x[n] = fma(-a.dt, x[n] * rmr[n], v[n]);  // Qtens = Qdp - dt*div
  }
  if (MODE == 3) {
    double s[16], lap[16];
    a.qtens.load(G, t.e, t.q, t.k, s);  // rspheremp * DSS(lap(Q))
    const double* T = G.T + (size_t)t.e * 48;
    laplace_wk(s, D, T, T + 16, T + 32, lap);
    const double cf = a.visc_coef * a.dp0[t.k];
    const double* rmp = G.rmp + (size_t)t.e * 16;
    TSE_UNROLL
    for (int n = 0; n < 16; ++n) x[n] = x[n] + cf * lap[n] * rmp[n];
  }
  {
    double c[16], rd[16];
    load16(a.pkg + 3 * a.pkg_stride + lp, c);
    load16(a.pkg + 4 * a.pkg_stride + lp, rd);
    limiter_optim_iter_full(x, c, rd, minp, maxp);
    TSE_UNROLL
    for (int n = 0; n < 16; ++n) x[n] = x[n] * c[n];  // spheremp * (x*dpmass)
  }
  store16(a.out + p * 16, x);
  a.mm.qmin[p] = minp;
  a.mm.qmax[p] = maxp;
}

// qdp_time_avg (prim_advection_mod.F90:645-662) fused with the pending DSS of the last stage
__global__ void __launch_bounds__(GPL* QPB) k_time_avg(Geo G, DssView np1, const double* __restrict__ q0, double rkstage,
                                                       double* __restrict__ out) {
  const ThreadPlane t = thread_plane(G, np1.Q);
  if (!t.valid) return;
  double v[16], w[16];
  np1.load(G, t.e, t.q, t.k, v);
  const size_t p = qplane(t.e, t.q, t.k, np1.Q) * 16;
  load16(q0 + p, w);
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) v[n] = (w[n] + (rkstage - 1.0) * v[n]) / rkstage;
  store16(out + p, v);
}

// materialise a pending DSS (needed only when the host asks for the field or before the remap)
__global__ void __launch_bounds__(GPL* QPB) k_resolve(Geo G, DssView in, double* __restrict__ out) {
  const ThreadPlane t = thread_plane(G, in.Q);
  if (!t.valid) return;
  double v[16];
  in.load(G, t.e, t.q, t.k, v);
  store16(out + qplane(t.e, t.q, t.k, in.Q) * 16, v);
}

// ---------------------------------------------------------------------------------------------
// host <-> device layout conversion.  stage: dense [ne_chunk][planes_per_elem][16] in the Fortran order of one element.
// ---------------------------------------------------------------------------------------------
// tracer field: host plane index within an element = k + NLEV*q
__global__ void k_qdp_relayout(double* __restrict__ dev, double* __restrict__ stage, const int* __restrict__ h2i, int e0, int ne,
                               int Q, int to_device) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // double2 index in the staging chunk
  const size_t total = (size_t)ne * Q * NLEV * 8;
  if (idx >= total) return;
  const int c = idx % 8;
  const size_t hp = idx / 8;
  const int k = hp % NLEV, q = (hp / NLEV) % Q, eh = hp / ((size_t)NLEV * Q);
  const int e = h2i[e0 + eh];
  double2* d = reinterpret_cast<double2*>(dev + qplane(e, q, k, Q) * 16) + c;
  double2* s = reinterpret_cast<double2*>(stage) + idx;
  if (to_device) *d = *s; else *s = *d;
}
// level field with ncomp components: host plane index = c + ncomp*k (derived%vn0(np,np,2,nlev)); host_nlev >= NLEV (eta_dot has nlev+1)
__global__ void k_level_relayout(double* __restrict__ dev, double* __restrict__ stage, const int* __restrict__ h2i, int ne, int ncomp,
                                 int host_nlev, int to_device) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)ne * ncomp * NLEV * 8;
  if (idx >= total) return;
  const int c8 = idx % 8;
  const size_t hp = idx / 8;
  const int cc = hp % ncomp, k = (hp / ncomp) % NLEV, eh = hp / ((size_t)ncomp * NLEV);
  const int e = h2i[eh];
  const size_t dp_ = (ncomp == 2 ? vplane(e, k, cc) : lplane(e, k));
  double2* d = reinterpret_cast<double2*>(dev + dp_ * 16) + c8;
  double2* s = reinterpret_cast<double2*>(stage + ((size_t)eh * ncomp * host_nlev + (size_t)k * ncomp + cc) * 16) + c8;
  if (to_device) *d = *s; else *s = *d;
}
// per-plane scalars (qmin/qmax): host [e][q][k]
__global__ void k_scalar_to_host(const double* __restrict__ dev, double* __restrict__ host, const int* __restrict__ h2i, int ne, int Q) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)ne * Q * NLEV) return;
  const int k = idx % NLEV, q = (idx / NLEV) % Q, eh = idx / ((size_t)NLEV * Q);
  host[idx] = dev[qplane(h2i[eh], q, k, Q)];
}

}  // namespace tse
