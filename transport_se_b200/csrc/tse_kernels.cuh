// CUDA kernels of the tracer-advection path (sm_100a, FP64): level-field kernels, halo packing, layout conversion and the
// plane-per-thread view used by the diagnostics.  The tracer-field kernels live in tse_pipe.cuh / tse_tile.cuh.
//
// The DSS (edgeVpack / bndry_exchangeV / edgeVunpack, edge_mod.F90:366-742) is never materialised as an edge buffer:
// a kernel that consumes a field produced "pre-DSS" gathers the neighbour nodes in the reference's unpack order while
// loading (DssView::load here, the shared-memory tile + halo in tse_pipe.cuh).
#pragma once
#include "tse_ops.cuh"

namespace tse {

constexpr int QPB = 4;  // tracers per CTA in the plane-per-thread kernels

struct Geo {
  const double* spheremp;   // [e][16]
  const double* rspheremp;  // [e][16]
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
// One CTA per two level chunks of a group (NKC is even): the 32 KB of vn0 and the group's metric terms go through shared
// memory with coalesced 16-byte accesses (rows padded to 144 bytes: a thread then reads its own 128-byte plane conflict-free),
// and so do the 16 KB of results.  (One plane per thread straight from global memory ran at 2.5 TB/s.)
constexpr int DV_THREADS = 2 * GPL, DV_ROW = 144, DV_MD = 64 * 8 + 16, DV_RM = 16 * 8 + 16;
constexpr int DV_SMEM = 4 * GPL * DV_ROW + GE * DV_MD + GE * DV_RM;
static_assert(NKC % 2 == 0, "k_divdp pairs level chunks");
__global__ void __launch_bounds__(DV_THREADS) k_divdp(Geo G, Dvv D, const double* __restrict__ vn0, double* __restrict__ divdp,
                                                      double* __restrict__ divdp_proj) {
  extern __shared__ __align__(16) unsigned char dv_smem[];
  unsigned char* const sv = dv_smem;                    // [chunk][component][plane] rows of vn0
  unsigned char* const smd = sv + 4 * GPL * DV_ROW;     // [element][4][16] metric terms
  unsigned char* const srm = smd + GE * DV_MD;          // [element][16] rmetdet
  const int t = threadIdx.x;
  const size_t c0 = (size_t)blockIdx.x * 2;             // first chunk: (g, kc) = (c0 / NKC, c0 % NKC)
  const int g = (int)(c0 / NKC), kc0 = (int)(c0 % NKC);
  {
    const double2* src = reinterpret_cast<const double2*>(vn0 + c0 * 2 * GPL * 16);
    TSE_UNROLL
    for (int j = 0; j < 16; ++j) {
      const int i = t + DV_THREADS * j;
      *reinterpret_cast<double2*>(sv + (i >> 3) * DV_ROW + (i & 7) * 16) = src[i];
    }
    const int elast = G.nelem - 1;
    for (int i = t; i < GE * 32; i += DV_THREADS) {
      const int el = i >> 5, c = i & 31, e = min(g * GE + el, elast);
      *reinterpret_cast<double2*>(smd + el * DV_MD + c * 16) = *reinterpret_cast<const double2*>(G.mD + (size_t)e * 64 + 2 * c);
    }
    if (t < GE * 8) {
      const int el = t >> 3, c = t & 7, e = min(g * GE + el, elast);
      *reinterpret_cast<double2*>(srm + el * DV_RM + c * 16) = *reinterpret_cast<const double2*>(G.rmr + (size_t)e * 16 + 2 * c);
    }
  }
  __syncthreads();
  const int cl = t / GPL, pl = t % GPL, el = pl / KC;
  unsigned char* const row1 = sv + ((cl * 2) * GPL + pl) * DV_ROW;
  const unsigned char* const row2 = row1 + GPL * DV_ROW;
  double g1[16], g2[16], r[16];
  TSE_UNROLL
  for (int c = 0; c < 8; ++c) {
    const double2 a1 = *reinterpret_cast<const double2*>(row1 + c * 16), a2 = *reinterpret_cast<const double2*>(row2 + c * 16);
    const unsigned char* m = smd + el * DV_MD + c * 16;
    const double2 m11 = *reinterpret_cast<const double2*>(m), m12 = *reinterpret_cast<const double2*>(m + 128);
    const double2 m21 = *reinterpret_cast<const double2*>(m + 256), m22 = *reinterpret_cast<const double2*>(m + 384);
    g1[2 * c] = m11.x * a1.x + m12.x * a2.x;
    g1[2 * c + 1] = m11.y * a1.y + m12.y * a2.y;
    g2[2 * c] = m21.x * a1.x + m22.x * a2.x;
    g2[2 * c + 1] = m21.y * a1.y + m22.y * a2.y;
  }
  div_contract(g1, g2, D, r);
  TSE_UNROLL
  for (int c = 0; c < 8; ++c) {
    const double2 rm = *reinterpret_cast<const double2*>(srm + el * DV_RM + c * 16);
    *reinterpret_cast<double2*>(row1 + c * 16) = make_double2(r[2 * c] * rm.x, r[2 * c + 1] * rm.y);  // (only this thread reads row1)
  }
  __syncthreads();
  double2* const o1 = reinterpret_cast<double2*>(divdp + c0 * GPL * 16);
  double2* const o2 = reinterpret_cast<double2*>(divdp_proj + c0 * GPL * 16);
  TSE_UNROLL
  for (int j = 0; j < 8; ++j) {
    const int i = t + DV_THREADS * j, orow = i >> 3;
    if (g * GE + (orow % GPL) / KC >= G.nelem) continue;  // padding elements of the last group
    const double2 v = *reinterpret_cast<const double2*>(sv + ((orow / GPL) * 2 * GPL + orow % GPL) * DV_ROW + (i & 7) * 16);
    o1[i] = v;
    o2[i] = v;
  }
  (void)kc0;
}

// ---------------------------------------------------------------------------------------------
// tracer-field kernels (one thread per (element, level, tracer) plane)
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// halo packing for the multi-GPU exchange (the send side of bndry_exchangeV, bndry_mod.F90:74-112): the boundary nodes
// that the reference's edgeVpack would place in the slab of a neighbour rank, gathered into a contiguous send buffer
// with the layout of the receiver's ghost array.
// ---------------------------------------------------------------------------------------------
// tracer field: send[(i*Q + q)*NLEV + k] = field(elem(i), q, k, node(i))
__global__ void __launch_bounds__(256) k_pack_tracer(const double* __restrict__ f, const int* __restrict__ send_src, int nsend, int Q,
                                                     double* __restrict__ send) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)nsend * Q * NLEV;
  if (idx >= total) return;
  const int k = idx % NLEV, q = (idx / NLEV) % Q, i = idx / ((size_t)NLEV * Q);
  const int code = send_src[i];
  send[idx] = f[qplane(code >> 4, q, k, Q) * 16 + (code & 15)];
}
// level field (mass weighted, as edgeVpack sees it after "DSSvar = spheremp*DSSvar"): send[i*NLEV + k]
__global__ void __launch_bounds__(256) k_pack_level(const double* __restrict__ f, const double* __restrict__ spheremp,
                                                    const int* __restrict__ send_src, int nsend, double* __restrict__ send) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)nsend * NLEV) return;
  const int k = idx % NLEV, i = idx / NLEV;
  const int code = send_src[i], e = code >> 4, nd = code & 15;
  send[idx] = spheremp[(size_t)e * 16 + nd] * f[lplane(e, k) * 16 + nd];
}
// element extrema for neighbor_minmax: send[((b*2 + {0,1})*Q + q)*NLEV + k] = qmin_loc / qmax_loc of element mm_elem(b)
__global__ void __launch_bounds__(256) k_pack_minmax(const double* __restrict__ lmin, const double* __restrict__ lmax,
                                                     const int* __restrict__ mm_elem, int nb, int Q, double* __restrict__ send) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)nb * Q * NLEV;
  if (idx >= total) return;
  const int k = idx % NLEV, q = (idx / NLEV) % Q, b = idx / ((size_t)NLEV * Q);
  const size_t p = qplane(mm_elem[b], q, k, Q);
  send[((size_t)(b * 2) * Q + q) * NLEV + k] = lmin[p];
  send[((size_t)(b * 2 + 1) * Q + q) * NLEV + k] = lmax[p];
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
