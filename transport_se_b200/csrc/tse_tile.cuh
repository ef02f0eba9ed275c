// Tiled tracer-field kernels (sm_100a): the production versions of the plane-per-thread kernels in tse_kernels.cuh.
//
// CTA = (group of 16 elements, chunk of 4 levels), QI*64 threads; it walks all tracers, QI at a time.  For each step the
// QI*8 KB tile [q][el][kk][16] (contiguous in HBM) is copied with coalesced 16-byte cp.async into shared memory (IN buffer),
// XOR-swizzled per 128-byte plane so that a thread can read "its" plane with conflict-free 128-bit loads.  One thread owns
// one plane: all 4x4 contractions, the limiter and the extrema are register-only.  As soon as every thread has pulled its
// plane (and its DSS neighbours) into registers the next tile is prefetched into the same IN buffer, overlapping the whole
// compute phase; results are staged in a second (OUT) buffer and written back with coalesced 16-byte stores.
//
// A warp holds 8/QI elements x 4 levels x QI tracers, so that the data-dependent limiter loop diverges as little as possible
// (the behaviour of a plane is mostly a property of its element).
//
// DSS (edgeVpack / bndry_exchangeV / edgeVunpack, edge_mod.F90:366-742) is fused into the load of the consumer: a field is
// stored "pre-DSS" (spheremp-weighted); neighbours inside the group are read straight from the tile in shared memory,
// neighbours outside the group (the patch perimeter, 68 nodes for a 4x4 patch) are fetched by 8-byte cp.async into a halo
// array next to the tile, and the sum runs in the reference's unpack order.  The rspheremp factor of the DSS is folded into
// the per-level package below, computed once per CTA and reused for all tracers.
//
// Per-(element, level) package (shared memory, tracer independent), with rX = rspheremp if the input is pre-DSS else 1:
//   dp_s = dp - rhs_mult*dt*divdp_proj, Vstar = vn0/dp_s, dp_star = dp_s - dt*divdp        (prim_advection_mod.F90:753,847-864)
//   U_c  = rX * metdet*(Dinv(c,1)*Vstar1 + Dinv(c,2)*Vstar2)      gv_c = U_c * S       (S = raw DSS sum)
//   E1   = spheremp*rX, E2 = dt*spheremp*rmetdet*rrearth          y = spheremp*Qtens = E1*S - E2*div
//   CL   = spheremp*dp_star  (the limiter's c), RDP = rX/dp_s     (Q = S*RDP for the extrema)
// The limiter works on y = c*x directly (limiter_y), so neither 1/dp_star nor the final spheremp multiply is needed.
#pragma once
#include "tse_kernels.cuh"

namespace tse {

#ifndef TSE_QI
#define TSE_QI 2
#endif
#ifndef TSE_MINB
#define TSE_MINB 2
#endif
constexpr int QI = TSE_QI;             // tracers per pipeline step
constexpr int TT = QI * GPL;           // threads per CTA (256 for QI = 4)
constexpr int TILE_BYTES = TT * 128;   // 32 KB for QI = 4
constexpr int HPRE = 512 / TT;         // halo entries per thread resolved before the tracer loop (covers hmax <= 128)
constexpr int QW = QI < 8 ? QI : 8;    // tracers per warp
constexpr int EPW = 8 / QW;            // elements per warp (1 for QI >= 8: the limiter's work is a property of the element)
static_assert(KC == 4 && (QI % QW) == 0 && (GE % EPW) == 0 && TT % 32 == 0 && TT <= 512, "warp mapping");
static_assert(EPW > 1 || (TT / 8) % GPL == 0 || GPL % (TT / 8) == 0, "swizzle");

// XOR swizzle of the 16-byte units of a 128-byte plane: distinct over the 8 planes a quarter-warp reads together
// (lanes 0..7 = 4 levels x 2 elements for EPW > 1, 4 levels x 2 tracers for EPW == 1)
__host__ __device__ constexpr int swz(int p) { return EPW > 1 ? (p & 7) : ((p & 3) | (((p / GPL) & 1) << 2)); }

enum TileOp { OP_MINMAX = 0, OP_STAGE1, OP_STAGE2, OP_STAGE3, OP_BIHARM_PRE, OP_TIME_AVG, OP_RESOLVE };

struct TileTables {
  const int* gsrc_t;    // [npad][NSLOT]: <0 none, [0,256) in-group (el<<4|node), >=256 halo entry (code-256)
  const int* halo_off;  // [ngroups+1]
  const int* halo_src;  // [halo_off[ngroups]]: >=0 (elem<<4|node), <=-2 ghost slot
  int hmax;             // max halo entries of a group
};

struct TileArgs {
  const double* src[2];    // input fields.  STAGE3: [0] = qtens, [1] = Qdp;  TIME_AVG: [0] = Qdp(n0), [1] = Qdp(np1)
  int pending[2];
  const double* ghost[2];
  double* out;
  const double *vn0, *dp, *divdp, *divdp_proj;
  double rhs_mult_dt, dt, visc_coef, rkstage;
  const double* dp0;
  double *qmin, *qmax, *qmin_loc, *qmax_loc;
  int Q;
  const int* glist;  // optional list of groups this launch covers (boundary groups first, interior groups while the halo is in flight)
};

constexpr int PP_BYTES = GPL * 128;  // per-plane package field (8 KB)
constexpr int EL_BYTES = GE * 128;   // per-element package field (2 KB)

// which package fields an op keeps in shared memory
struct TileCfg {
  int npp, nel, has_out;
  int U1, U2, CL, RDP, RC;        // per-plane field slots
  int E1, E2, RSPH, T11, T12, T22;  // per-element field slots
};
__host__ __device__ constexpr TileCfg tile_cfg(int op) {
  return op == OP_STAGE1 ? TileCfg{4, 2, 1, 0, 1, 2, -1, 3, 0, 1, -1, -1, -1, -1}
       : op == OP_STAGE2 ? TileCfg{5, 2, 1, 0, 1, 2, 3, 4, 0, 1, -1, -1, -1, -1}
       : op == OP_STAGE3 ? TileCfg{4, 6, 1, 0, 1, 2, -1, 3, 0, 1, 2, 3, 4, 5}
       : op == OP_MINMAX ? TileCfg{1, 0, 0, -1, -1, -1, 0, -1, -1, -1, -1, -1, -1, -1}
       : op == OP_BIHARM_PRE ? TileCfg{1, 3, 1, -1, -1, -1, 0, -1, -1, -1, -1, 0, 1, 2}
                             : TileCfg{0, 1, 1, -1, -1, -1, -1, -1, -1, -1, 0, -1, -1, -1};
}
__host__ __device__ constexpr int tile_in_bytes(int hmax) { return TILE_BYTES + QI * hmax * KC * 8 + 16; }
constexpr int NBUF = 2;  // IN buffers: two tiles are in flight while a third is being processed from registers
__host__ __device__ constexpr int tile_smem_bytes(int op, int hmax) {
  return NBUF * tile_in_bytes(hmax) + (tile_cfg(op).has_out ? TILE_BYTES : 0) + tile_cfg(op).npp * PP_BYTES + tile_cfg(op).nel * EL_BYTES;
}

__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;\n" ::); }

__device__ __forceinline__ double2 lds128(const unsigned char* base, int off) { return *reinterpret_cast<const double2*>(base + off); }
__device__ __forceinline__ double lds64(const unsigned char* base, int off) { return *reinterpret_cast<const double*>(base + off); }

// volatile shared-memory load: keeps the limiter's c out of registers (re-read from the package each pass)
__device__ __forceinline__ double2 lds128v(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// limiter_optim_iter_full (prim_advection_mod.F90:976-1094).  On entry y = c*x (mass contributions, c = sphweights*dpmass);
// c and rc = 1/c are read from the per-plane package in shared memory (cbase/rcbase = shared address of chunk 0 of this
// plane, chunk stride GPL*16 bytes).  Sums use 4 interleaved partial accumulators (fixed order, identical on every GPU count).
//
// Fast path: mass = sum(y) and the min/max relaxation (:1016-1029) need no x; if no x = y*rc lies outside [minp, maxp] the
// reference's first sweep finds addmass = 0 and leaves (:1047), so y is returned untouched.
// Slow path (x in place of y): sweep 1 clips against both bounds.  From then on the direction is fixed: redistributing
// addmass > 0 raises nodes below maxp, so later sweeps can only find nodes above maxp and addmass stays >= 0 (and the mirror
// image for addmass < 0); the lower-bound test of the reference's sweeps 2..15 is then never taken.  Working on z = -x,
// bound -minp for the downward case (negation is exact) leaves one code path, whose sweep fuses "add the increment"
// (:1052-1078 of sweep i), "clip" (:1037-1045 of sweep i+1) and the next weightssum.
__device__ __forceinline__ void limiter_y(double (&y)[16], unsigned cbase, unsigned rcbase, double sumc, double& minp, double& maxp) {
  const double tol_limiter = (double)5e-14f;
  if (sumc <= 0.0) return;
  double mass;
  {
    double m0 = y[0], m1 = y[1], m2 = y[2], m3 = y[3];
    TSE_UNROLL
    for (int n = 4; n < 16; n += 4) {
      m0 += y[n];
      m1 += y[n + 1];
      m2 += y[n + 2];
      m3 += y[n + 3];
    }
    mass = (m0 + m1) + (m2 + m3);
  }
  if (mass < minp * sumc) minp = mass / sumc;
  if (mass > maxp * sumc) maxp = mass / sumc;
  bool viol = false;
  TSE_UNROLL
  for (int cc = 0; cc < 8; ++cc) {
    const double2 r = lds128v(rcbase + cc * GPL * 16);
    const double x0 = y[2 * cc] * r.x, x1 = y[2 * cc + 1] * r.y;
    viol |= (x0 > maxp) | (x0 < minp) | (x1 > maxp) | (x1 < minp);
  }
  if (!viol) return;

  const double thresh = tol_limiter * fabs(mass);
  TSE_UNROLL
  for (int cc = 0; cc < 8; ++cc) {
    const double2 r = lds128v(rcbase + cc * GPL * 16);
    y[2 * cc] *= r.x;
    y[2 * cc + 1] *= r.y;
  }
  double am;
  {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    TSE_UNROLL
    for (int cc = 0; cc < 8; cc += 2) {
      const double2 ca = lds128v(cbase + cc * GPL * 16), cb = lds128v(cbase + (cc + 1) * GPL * 16);
      const int n = 2 * cc;
      double t;
      t = dmin(dmax(y[n], minp), maxp);         a0 = fma(y[n] - t, ca.x, a0);     y[n] = t;
      t = dmin(dmax(y[n + 1], minp), maxp);     a1 = fma(y[n + 1] - t, ca.y, a1); y[n + 1] = t;
      t = dmin(dmax(y[n + 2], minp), maxp);     a2 = fma(y[n + 2] - t, cb.x, a2); y[n + 2] = t;
      t = dmin(dmax(y[n + 3], minp), maxp);     a3 = fma(y[n + 3] - t, cb.y, a3); y[n + 3] = t;
    }
    am = (a0 + a1) + (a2 + a3);
  }
  if (fabs(am) > thresh) {
    const bool up = am > 0.0;
    const double bz = up ? maxp : -minp;
    if (!up) {
      am = -am;
      TSE_UNROLL
      for (int n = 0; n < 16; ++n) y[n] = -y[n];
    }
    double wsum;
    {
      double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
      TSE_UNROLL
      for (int cc = 0; cc < 8; cc += 2) {
        const double2 ca = lds128v(cbase + cc * GPL * 16), cb = lds128v(cbase + (cc + 1) * GPL * 16);
        const int n = 2 * cc;
        if (y[n] < bz) w0 += ca.x;
        if (y[n + 1] < bz) w1 += ca.y;
        if (y[n + 2] < bz) w2 += cb.x;
        if (y[n + 3] < bz) w3 += cb.y;
      }
      wsum = (w0 + w1) + (w2 + w3);
    }
#pragma unroll 1
    for (int iter = 1; iter <= NPSQ - 1; ++iter) {
      const double inc = am / wsum;
      if (iter == NPSQ - 1) {  // the reference's last sweep redistributes without a further clip
        TSE_UNROLL
        for (int n = 0; n < 16; ++n)
          if (y[n] < bz) y[n] += inc;
        break;
      }
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
      TSE_UNROLL
      for (int cc = 0; cc < 8; cc += 2) {
        const double2 ca = lds128v(cbase + cc * GPL * 16), cb = lds128v(cbase + (cc + 1) * GPL * 16);
        const double cv[4] = {ca.x, ca.y, cb.x, cb.y};
        TSE_UNROLL
        for (int u = 0; u < 4; ++u) {
          const int n = 2 * cc + u;
          double z = y[n];
          if (z < bz) z += inc;
          const double d = z - bz;
          double ad = 0.0, wd = 0.0;
          if (d > 0.0) { ad = d; z = bz; }
          if (d < 0.0) wd = cv[u];
          y[n] = z;
          if (u == 0) { a0 = fma(ad, cv[u], a0); w0 += wd; }
          if (u == 1) { a1 = fma(ad, cv[u], a1); w1 += wd; }
          if (u == 2) { a2 = fma(ad, cv[u], a2); w2 += wd; }
          if (u == 3) { a3 = fma(ad, cv[u], a3); w3 += wd; }
        }
      }
      am = (a0 + a1) + (a2 + a3);
      wsum = (w0 + w1) + (w2 + w3);
      if (am <= thresh) break;
    }
    if (!up) {
      TSE_UNROLL
      for (int n = 0; n < 16; ++n) y[n] = -y[n];
    }
  }
  TSE_UNROLL
  for (int cc = 0; cc < 8; ++cc) {
    const double2 c = lds128v(cbase + cc * GPL * 16);
    y[2 * cc] *= c.x;
    y[2 * cc + 1] *= c.y;
  }
}

// y(a,b) = sum_i Dvv(i,a) g1(i,b) + sum_i Dvv(i,b) g2(a,i) with g_c = U_c*S, evaluated row by row / column pair by column pair
// so that only S, y and 8 temporaries are live (div_contract needs S, g1, g2 and y at once)
__device__ __forceinline__ void flux_div(const double (&S)[16], const unsigned char* u1, const unsigned char* u2, int pl, const Dvv& D,
                                         double (&y)[16]) {
  TSE_UNROLL
  for (int b = 0; b < 4; ++b) {
    const double2 ua = lds128(u1, ((2 * b) * GPL + pl) * 16), ub = lds128(u1, ((2 * b + 1) * GPL + pl) * 16);
    const double g0 = ua.x * S[4 * b], g1 = ua.y * S[4 * b + 1], g2 = ub.x * S[4 * b + 2], g3 = ub.y * S[4 * b + 3];
    y[4 * b + 0] = fma(D.d[3 + 0], g3, fma(D.d[2 + 0], g2, fma(D.d[1 + 0], g1, D.d[0 + 0] * g0)));
    y[4 * b + 1] = fma(D.d[3 + 4], g3, fma(D.d[2 + 4], g2, D.d[0 + 4] * g0));   // Dvv(1,1) = 0
    y[4 * b + 2] = fma(D.d[3 + 8], g3, fma(D.d[1 + 8], g1, D.d[0 + 8] * g0));   // Dvv(2,2) = 0
    y[4 * b + 3] = fma(D.d[3 + 12], g3, fma(D.d[2 + 12], g2, fma(D.d[1 + 12], g1, D.d[0 + 12] * g0)));
  }
  TSE_UNROLL
  for (int h = 0; h < 2; ++h) {  // columns a = 2h, 2h+1: nodes a + 4 i live in chunks h + 2 i
    double ge[4], go[4];
    TSE_UNROLL
    for (int i = 0; i < 4; ++i) {
      const double2 u = lds128(u2, ((h + 2 * i) * GPL + pl) * 16);
      ge[i] = u.x * S[2 * h + 4 * i];
      go[i] = u.y * S[2 * h + 1 + 4 * i];
    }
    TSE_UNROLL
    for (int b = 0; b < 4; ++b) {
      double se = 0.0, so = 0.0;
      TSE_UNROLL
      for (int i = 0; i < 4; ++i) {
        if (!(i == b && (b == 1 || b == 2))) {
          se = fma(D.d[i + 4 * b], ge[i], se);
          so = fma(D.d[i + 4 * b], go[i], so);
        }
      }
      y[2 * h + 4 * b] += se;
      y[2 * h + 1 + 4 * b] += so;
    }
  }
}

// laplace_sphere_wk with the per-element tensor T read from the element-level package
__device__ __forceinline__ void laplace_wk_el(const double (&s)[16], const Dvv& D, const unsigned char* t11, const unsigned char* t12,
                                              const unsigned char* t22, int el, double (&lap)[16]) {
  double w1[16], w2[16];
  {
    double d1[16], d2[16];
    grad_raw(s, D, d1, d2);
    TSE_UNROLL
    for (int c = 0; c < 8; ++c) {
      const double2 a = lds128(t11, (c * GE + el) * 16);
      const double2 b = lds128(t12, (c * GE + el) * 16);
      const double2 cc = lds128(t22, (c * GE + el) * 16);
      w1[2 * c] = a.x * d1[2 * c] + b.x * d2[2 * c];
      w2[2 * c] = b.x * d1[2 * c] + cc.x * d2[2 * c];
      w1[2 * c + 1] = a.y * d1[2 * c + 1] + b.y * d2[2 * c + 1];
      w2[2 * c + 1] = b.y * d1[2 * c + 1] + cc.y * d2[2 * c + 1];
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

template <int OP>
__global__ void __launch_bounds__(TT, TSE_MINB) k_tile(Geo G, Dvv D, TileTables tb, TileArgs a) {
  constexpr TileCfg cfg = tile_cfg(OP);
  constexpr bool kStage = (OP == OP_STAGE1 || OP == OP_STAGE2 || OP == OP_STAGE3);
  constexpr int NIN = (OP == OP_STAGE3 || OP == OP_TIME_AVG) ? 2 : 1;
  constexpr bool kHasOut = cfg.has_out != 0;
  extern __shared__ __align__(16) unsigned char smem[];
  const int IN_BYTES = tile_in_bytes(tb.hmax);
  unsigned char* const outb = smem + NBUF * IN_BYTES;
  unsigned char* const pp = outb + (kHasOut ? TILE_BYTES : 0);
  unsigned char* const elb = pp + cfg.npp * PP_BYTES;
  const int ZERO_OFF = TILE_BYTES + QI * tb.hmax * KC * 8;

  const int t = threadIdx.x;
  const int g = a.glist ? a.glist[blockIdx.x / NKC] : blockIdx.x / NKC, kc = blockIdx.x % NKC;
  // warp = EPW elements x 4 levels x QW tracers
  const int w = t >> 5, lane = t & 31;
  const int kk = lane & 3, el = EPW * (w % (GE / EPW)) + ((lane >> 2) % EPW), qi = QW * (w / (GE / EPW)) + (lane >> 2) / EPW;
  const int pl = el * KC + kk;   // plane within one tracer's tile
  const int p = qi * GPL + pl;   // plane within the QI-tracer tile
  const int e = g * GE + el, k = kc * KC + kk;
  const bool evalid = e < G.nelem;
  const int Q = a.Q;
  const int hoff = tb.halo_off[g], H = tb.halo_off[g + 1] - hoff;

  // ---- level package -------------------------------------------------------------------------------------------
  const bool main_pending = (OP == OP_STAGE3) ? (a.pending[1] != 0) : (a.pending[0] != 0);
  if (t < NBUF) *reinterpret_cast<double*>(smem + t * IN_BYTES + ZERO_OFF) = 0.0;
  if (cfg.nel > 0 && t < GE * 8) {
    // element-level fields: thread -> (element t>>3, nodes 2*(t&7), +1)
    const int pe = t >> 3, c = t & 7, ee = g * GE + pe;
    double2 e1 = make_double2(0, 0), e2 = e1, rs = e1, t11 = e1, t12 = e1, t22 = e1;
    if (ee < G.nelem) {
      const size_t b = (size_t)ee * 16 + 2 * c;
      const double2 sp = *reinterpret_cast<const double2*>(G.spheremp + b);
      rs = *reinterpret_cast<const double2*>(G.rspheremp + b);
      const double2 rm = *reinterpret_cast<const double2*>(G.rmr + b);
      const double rx0 = main_pending ? rs.x : 1.0, rx1 = main_pending ? rs.y : 1.0;
      e1 = make_double2(sp.x * rx0, sp.y * rx1);
      e2 = make_double2(a.dt * (sp.x * rm.x), a.dt * (sp.y * rm.y));
      if (cfg.T11 >= 0) {
        const double* T = G.T + (size_t)ee * 48 + 2 * c;
        t11 = *reinterpret_cast<const double2*>(T);
        t12 = *reinterpret_cast<const double2*>(T + 16);
        t22 = *reinterpret_cast<const double2*>(T + 32);
      }
    }
    const int off = (c * GE + pe) * 16;
    if (cfg.E1 >= 0) *reinterpret_cast<double2*>(elb + cfg.E1 * EL_BYTES + off) = e1;
    if (cfg.E2 >= 0) *reinterpret_cast<double2*>(elb + cfg.E2 * EL_BYTES + off) = e2;
    if (cfg.RSPH >= 0) *reinterpret_cast<double2*>(elb + cfg.RSPH * EL_BYTES + off) = rs;
    if (cfg.T11 >= 0) {
      *reinterpret_cast<double2*>(elb + cfg.T11 * EL_BYTES + off) = t11;
      *reinterpret_cast<double2*>(elb + cfg.T12 * EL_BYTES + off) = t12;
      *reinterpret_cast<double2*>(elb + cfg.T22 * EL_BYTES + off) = t22;
    }
  }
  if (cfg.npp > 0) {
    // per-plane fields: thread -> (plane t / PARTS, 8 / PARTS consecutive 16-byte chunks)
    constexpr int PARTS = TT / GPL;
    const int ppl = t / PARTS, part = t % PARTS;
    const int pe = g * GE + ppl / KC, pk = kc * KC + ppl % KC;
    TSE_UNROLL
    for (int cc = 0; cc < 8 / PARTS; ++cc) {
      const int c = part * (8 / PARTS) + cc, n = 2 * c;
      double2 u1 = make_double2(0, 0), u2 = u1, cl = make_double2(1, 1), rd = make_double2(1, 1), rcl = make_double2(1, 1);
      if (pe < G.nelem) {
        const size_t lp = lplane(pe, pk) * 16 + n, gb = (size_t)pe * 16 + n;
        const double2 dpv = *reinterpret_cast<const double2*>(a.dp + lp);
        const double2 dj = *reinterpret_cast<const double2*>(a.divdp_proj + lp);
        const double2 rs = *reinterpret_cast<const double2*>(G.rspheremp + gb);
        const double rx0 = main_pending ? rs.x : 1.0, rx1 = main_pending ? rs.y : 1.0;
        const double dps0 = dpv.x - a.rhs_mult_dt * dj.x, dps1 = dpv.y - a.rhs_mult_dt * dj.y;
        const double r0 = 1.0 / dps0, r1 = 1.0 / dps1;
        rd = make_double2(r0 * rx0, r1 * rx1);
        if (kStage) {
          const double2 dd = *reinterpret_cast<const double2*>(a.divdp + lp);
          const double2 v1 = *reinterpret_cast<const double2*>(a.vn0 + vplane(pe, pk, 0) * 16 + n);
          const double2 v2 = *reinterpret_cast<const double2*>(a.vn0 + vplane(pe, pk, 1) * 16 + n);
          const double2 sp = *reinterpret_cast<const double2*>(G.spheremp + gb);
          const double* mD = G.mD + (size_t)pe * 64 + n;
          const double2 m11 = *reinterpret_cast<const double2*>(mD), m12 = *reinterpret_cast<const double2*>(mD + 16);
          const double2 m21 = *reinterpret_cast<const double2*>(mD + 32), m22 = *reinterpret_cast<const double2*>(mD + 48);
          const double vs10 = v1.x * r0, vs11 = v1.y * r1, vs20 = v2.x * r0, vs21 = v2.y * r1;
          u1 = make_double2((m11.x * vs10 + m12.x * vs20) * rx0, (m11.y * vs11 + m12.y * vs21) * rx1);
          u2 = make_double2((m21.x * vs10 + m22.x * vs20) * rx0, (m21.y * vs11 + m22.y * vs21) * rx1);
          cl = make_double2(sp.x * (dps0 - a.dt * dd.x), sp.y * (dps1 - a.dt * dd.y));
          rcl = make_double2(1.0 / cl.x, 1.0 / cl.y);
        }
      }
      const int off = (c * GPL + ppl) * 16;
      if (cfg.U1 >= 0) *reinterpret_cast<double2*>(pp + cfg.U1 * PP_BYTES + off) = u1;
      if (cfg.U2 >= 0) *reinterpret_cast<double2*>(pp + cfg.U2 * PP_BYTES + off) = u2;
      if (cfg.CL >= 0) *reinterpret_cast<double2*>(pp + cfg.CL * PP_BYTES + off) = cl;
      if (cfg.RDP >= 0) *reinterpret_cast<double2*>(pp + cfg.RDP * PP_BYTES + off) = rd;
      if (cfg.RC >= 0) *reinterpret_cast<double2*>(pp + cfg.RC * PP_BYTES + off) = rcl;
    }
  }

  // ---- per-thread DSS gather offsets (bytes inside the IN buffer), two 16-bit offsets per register ----------------
  unsigned goff[NSLOT / 2];
  {
    const int* gs = tb.gsrc_t + (size_t)(evalid ? e : 0) * NSLOT;
    TSE_UNROLL
    for (int s = 0; s < NSLOT; ++s) {
      const int code = evalid ? gs[s] : -1;
      int off = ZERO_OFF;
      if (code >= 256) off = TILE_BYTES + ((qi * tb.hmax + (code - 256)) * KC + kk) * 8;
      else if (code >= 0) {
        const int p2 = qi * GPL + (code >> 4) * KC + kk, node = code & 15;
        off = p2 * 128 + ((((node >> 1) ^ swz(p2))) << 4) + (node & 1) * 8;
      }
      // IN buffer is < 512 KB / 8: store offsets in units of 8 bytes
      if (s & 1) goff[s >> 1] |= (unsigned)(off >> 3) << 16;
      else goff[s >> 1] = (unsigned)(off >> 3);
    }
  }
  auto gofs = [&](int s) -> int { return (int)(((s & 1) ? (goff[s >> 1] >> 16) : (goff[s >> 1] & 0xffffu)) << 3); };

  // ---- halo entries handled by this thread: entry idx -> (h = idx % H, kk2 = idx / H), h fastest for coalescing --------
  const int nhalo = H * KC;
  long long hsrc[HPRE];  // double index of the source for tracer 0 (>= 0: in field; < 0: -(index in ghost array) - 1)
  int hdst[HPRE];        // byte offset in the IN buffer for tracer slot 0
  TSE_UNROLL
  for (int r = 0; r < HPRE; ++r) {
    const int idx = t + r * TT;
    hsrc[r] = 0;
    hdst[r] = -1;
    if (idx < nhalo) {
      const int h = idx % H, kk2 = idx / H;
      const int code = tb.halo_src[hoff + h];
      const int kq = kc * KC + kk2;
      hdst[r] = TILE_BYTES + (h * KC + kk2) * 8;
      if (code >= 0) hsrc[r] = (long long)(qplane(code >> 4, 0, kq, Q) * 16 + (code & 15));
      else hsrc[r] = -((long long)(-code - 2) * Q * NLEV + kq) - 1;
    }
  }

  const unsigned smem_u32 = (unsigned)__cvta_generic_to_shared(smem);
  const int nit = (Q + QI - 1) / QI;
  const int nitems = nit * NIN;
  const size_t cta_base = ((size_t)g * NKC + kc) * Q * GPL * 16;  // doubles
  // own-plane chunk offsets
  const int own_base = p * 128, own_sw = swz(p);
  // chunk i = r*TT + t of a tile -> plane r*TT/8 + (t>>3), 16-byte unit t&7
  auto cp_dst = [&](int r) -> int {
    const int pi = r * (TT / 8) + (t >> 3);
    return pi * 128 + (((t & 7) ^ swz(pi)) << 4);
  };
  const bool group_full = (g * GE + GE <= G.nelem);

  auto issue = [&](int j) {
    const int it = j / NIN, which = j % NIN;
    const double* src = a.src[which];
    const int q0 = it * QI, nq = min(QI, Q - q0);
    const double* tsrc = src + cta_base + (size_t)q0 * GPL * 16 + t * 2;
    const unsigned sb = smem_u32 + (j % NBUF) * IN_BYTES;
    if (nq == QI) {
      TSE_UNROLL
      for (int r = 0; r < 8; ++r) cp_async16(sb + cp_dst(r), tsrc + r * (TT * 2));
    } else {
      TSE_UNROLL
      for (int r = 0; r < 8; ++r)
        if (r * (TT / 8) + (t >> 3) < nq * GPL) cp_async16(sb + cp_dst(r), tsrc + r * (TT * 2));
    }
    if (a.pending[which]) {
      TSE_UNROLL
      for (int r = 0; r < HPRE; ++r) {
        if (hdst[r] >= 0) {
          for (int qi2 = 0; qi2 < nq; ++qi2) {
            const double* gp = hsrc[r] >= 0 ? src + hsrc[r] + (size_t)(q0 + qi2) * GPL * 16
                                            : a.ghost[which] + (-(hsrc[r] + 1)) + (size_t)(q0 + qi2) * NLEV;
            cp_async8(sb + hdst[r] + qi2 * tb.hmax * KC * 8, gp);
          }
        }
      }
      for (int idx = t + HPRE * TT; idx < nhalo; idx += TT) {  // groups with more than HPRE*TT/KC halo nodes (irregular patches)
        const int h = idx % H, kk2 = idx / H;
        const int code = tb.halo_src[hoff + h];
        const int kq = kc * KC + kk2;
        for (int qi2 = 0; qi2 < nq; ++qi2) {
          const double* gp = (code >= 0) ? src + (qplane(code >> 4, q0 + qi2, kq, Q) * 16 + (code & 15))
                                         : a.ghost[which] + ((size_t)(-code - 2) * Q + q0 + qi2) * NLEV + kq;
          cp_async8(sb + TILE_BYTES + ((qi2 * tb.hmax + h) * KC + kk2) * 8, gp);
        }
      }
    }
    cp_async_commit();
  };

  // q-invariant per-thread values
  double sumc = 0.0;
  __syncthreads();  // package visible
  if (kStage) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    TSE_UNROLL
    for (int c = 0; c < 8; c += 2) {
      const double2 x0 = lds128(pp + cfg.CL * PP_BYTES, (c * GPL + pl) * 16);
      const double2 x1 = lds128(pp + cfg.CL * PP_BYTES, ((c + 1) * GPL + pl) * 16);
      s0 += x0.x; s1 += x0.y; s2 += x1.x; s3 += x1.y;
    }
    sumc = (s0 + s1) + (s2 + s3);
  }
  const double cf = (OP == OP_STAGE3) ? a.visc_coef * a.dp0[k] : 0.0;

  double keep[16];  // STAGE3: cf*lap of the first item; TIME_AVG: Qdp(n0)
  // limiter bounds of this thread's plane, fetched one tracer step of the loop ahead (a global load the math depends on)
  const size_t pidx0 = (((size_t)g * NKC + kc) * Q + qi) * GPL + pl;
  double minp_n = 0.0, maxp_n = 0.0;
  if (kStage && evalid && qi < Q) {
    minp_n = a.qmin[pidx0];
    maxp_n = a.qmax[pidx0];
  }
  issue(0);
  if (nitems > 1) issue(1);
  for (int j = 0; j < nitems; ++j) {
    if (j + 1 < nitems) cp_async_wait_1(); else cp_async_wait_all();
    __syncthreads();  // IN[j % NBUF] holds item j; every thread is past the copy-out of the previous item
    const unsigned char* inb = smem + (j % NBUF) * IN_BYTES;
    const int it = j / NIN, which = j % NIN;
    const int q = it * QI + qi;
    const bool valid = evalid && q < Q;
    const size_t pidx = (((size_t)g * NKC + kc) * Q + q) * GPL + pl;  // global plane index

    double S[16];
    TSE_UNROLL
    for (int c = 0; c < 8; ++c) {
      const double2 v = lds128(inb, own_base + ((c ^ own_sw) << 4));
      S[2 * c] = v.x;
      S[2 * c + 1] = v.y;
    }
    if (a.pending[which]) {  // DSS in the reference's unpack order: S, E, N, W edges, then SW, SE, NE, NW corners
      TSE_UNROLL
      for (int i = 0; i < 4; ++i) S[i] += lds64(inb, gofs(i));
      TSE_UNROLL
      for (int i = 0; i < 4; ++i) S[3 + 4 * i] += lds64(inb, gofs(4 + i));
      TSE_UNROLL
      for (int i = 0; i < 4; ++i) S[12 + i] += lds64(inb, gofs(8 + i));
      TSE_UNROLL
      for (int i = 0; i < 4; ++i) S[4 * i] += lds64(inb, gofs(12 + i));
      S[0] += lds64(inb, gofs(16));
      S[3] += lds64(inb, gofs(17));
      S[15] += lds64(inb, gofs(18));
      S[12] += lds64(inb, gofs(19));
    }
    __syncthreads();  // all reads of this IN buffer done: refill it with the item after next, overlapping the compute below
    if (j + NBUF < nitems) issue(j + NBUF);

    const bool last_of_iter = (which == NIN - 1);
    if (valid) {
      if (OP == OP_MINMAX || OP == OP_BIHARM_PRE) {
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) {
          const double2 rd = lds128(pp + cfg.RDP * PP_BYTES, (c * GPL + pl) * 16);
          S[2 * c] *= rd.x;
          S[2 * c + 1] *= rd.y;
        }
        double mn0 = dmin(S[0], S[1]), mx0 = dmax(S[0], S[1]), mn1 = dmin(S[2], S[3]), mx1 = dmax(S[2], S[3]);
        TSE_UNROLL
        for (int n = 4; n < 16; n += 4) {
          mn0 = dmin(mn0, dmin(S[n], S[n + 1]));
          mx0 = dmax(mx0, dmax(S[n], S[n + 1]));
          mn1 = dmin(mn1, dmin(S[n + 2], S[n + 3]));
          mx1 = dmax(mx1, dmax(S[n + 2], S[n + 3]));
        }
        a.qmin_loc[pidx] = dmin(mn0, mn1);
        a.qmax_loc[pidx] = dmax(mx0, mx1);
        if (OP == OP_BIHARM_PRE) {
          double lap[16];
          laplace_wk_el(S, D, elb + cfg.T11 * EL_BYTES, elb + cfg.T12 * EL_BYTES, elb + cfg.T22 * EL_BYTES, el, lap);
          TSE_UNROLL
          for (int n = 0; n < 16; ++n) S[n] = lap[n];
        }
      } else if (OP == OP_RESOLVE) {
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) {
          const double2 rs = lds128(elb + cfg.RSPH * EL_BYTES, (c * GE + el) * 16);
          S[2 * c] *= rs.x;
          S[2 * c + 1] *= rs.y;
        }
      } else if (OP == OP_TIME_AVG) {
        if (which == 0) {
          TSE_UNROLL
          for (int n = 0; n < 16; ++n) keep[n] = S[n];
        } else {
          TSE_UNROLL
          for (int c = 0; c < 8; ++c) {
            double2 rs = lds128(elb + cfg.RSPH * EL_BYTES, (c * GE + el) * 16);
            if (!a.pending[1]) rs = make_double2(1.0, 1.0);
            S[2 * c] = (keep[2 * c] + (a.rkstage - 1.0) * (rs.x * S[2 * c])) / a.rkstage;
            S[2 * c + 1] = (keep[2 * c + 1] + (a.rkstage - 1.0) * (rs.y * S[2 * c + 1])) / a.rkstage;
          }
        }
      } else if (OP == OP_STAGE3 && which == 0) {
        // second half of biharmonic_wk_scalar_minmax: lap(rspheremp*DSS(qtens)); Qtens_biharmonic*spheremp = cf*lap
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) {
          const double2 rs = lds128(elb + cfg.RSPH * EL_BYTES, (c * GE + el) * 16);
          S[2 * c] *= rs.x;
          S[2 * c + 1] *= rs.y;
        }
        double lap[16];
        laplace_wk_el(S, D, elb + cfg.T11 * EL_BYTES, elb + cfg.T12 * EL_BYTES, elb + cfg.T22 * EL_BYTES, el, lap);
        TSE_UNROLL
        for (int n = 0; n < 16; ++n) keep[n] = cf * lap[n];
      } else if (kStage) {
        double minp = minp_n, maxp = maxp_n;
        if (evalid && q + QI < Q) {
          minp_n = a.qmin[pidx + (size_t)QI * GPL];
          maxp_n = a.qmax[pidx + (size_t)QI * GPL];
        }
        if (OP == OP_STAGE2) {
          double mn0 = 1e300, mx0 = -1e300, mn1 = 1e300, mx1 = -1e300;
          TSE_UNROLL
          for (int c = 0; c < 8; ++c) {
            const double2 rd = lds128(pp + cfg.RDP * PP_BYTES, (c * GPL + pl) * 16);
            const double q0v = S[2 * c] * rd.x, q1v = S[2 * c + 1] * rd.y;
            mn0 = dmin(mn0, q0v);
            mx0 = dmax(mx0, q0v);
            mn1 = dmin(mn1, q1v);
            mx1 = dmax(mx1, q1v);
          }
          minp = dmin(minp, dmin(mn0, mn1));
          maxp = dmax(maxp, dmax(mx0, mx1));
        }
        double y[16];
        asm volatile("" ::: "memory");
#ifdef TSE_SKIP_DIV
        TSE_UNROLL
        for (int n = 0; n < 16; ++n) y[n] = S[n];
#else
        flux_div(S, pp + cfg.U1 * PP_BYTES, pp + cfg.U2 * PP_BYTES, pl, D, y);
#endif
        asm volatile("" ::: "memory");
        const unsigned e1a = smem_u32 + (unsigned)(elb - smem) + (cfg.E1 < 0 ? 0 : cfg.E1) * EL_BYTES + el * 16;
        const unsigned e2a = smem_u32 + (unsigned)(elb - smem) + (cfg.E2 < 0 ? 0 : cfg.E2) * EL_BYTES + el * 16;
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) {
          const double2 e1 = lds128v(e1a + c * GE * 16);
          const double2 e2 = lds128v(e2a + c * GE * 16);
          y[2 * c] = fma(-e2.x, y[2 * c], e1.x * S[2 * c]);
          y[2 * c + 1] = fma(-e2.y, y[2 * c + 1], e1.y * S[2 * c + 1]);
          if (OP == OP_STAGE3) {
            y[2 * c] += keep[2 * c];
            y[2 * c + 1] += keep[2 * c + 1];
          }
        }
        asm volatile("" ::: "memory");
#ifndef TSE_SKIP_LIMITER
        limiter_y(y, smem_u32 + (unsigned)(pp - smem) + (cfg.CL < 0 ? 0 : cfg.CL) * PP_BYTES + pl * 16,
                  smem_u32 + (unsigned)(pp - smem) + (cfg.RC < 0 ? 0 : cfg.RC) * PP_BYTES + pl * 16, sumc, minp, maxp);
#endif
        asm volatile("" ::: "memory");
        a.qmin[pidx] = minp;
        a.qmax[pidx] = maxp;
        TSE_UNROLL
        for (int n = 0; n < 16; ++n) S[n] = y[n];
      }
    }

    if (kHasOut && last_of_iter) {
      if (valid) {
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) *reinterpret_cast<double2*>(outb + own_base + ((c ^ own_sw) << 4)) = make_double2(S[2 * c], S[2 * c + 1]);
      }
      __syncthreads();
      const int q0 = it * QI, nq = min(QI, Q - q0);
      double* tdst = a.out + cta_base + (size_t)q0 * GPL * 16 + t * 2;
      if (nq == QI && group_full) {
        TSE_UNROLL
        for (int r = 0; r < 8; ++r) *reinterpret_cast<double2*>(tdst + r * (TT * 2)) = lds128(outb, cp_dst(r));
      } else {
        TSE_UNROLL
        for (int r = 0; r < 8; ++r) {
          const int pi = r * (TT / 8) + (t >> 3);
          if (pi < nq * GPL && g * GE + ((pi / KC) % GE) < G.nelem)
            *reinterpret_cast<double2*>(tdst + r * (TT * 2)) = lds128(outb, cp_dst(r));
        }
      }
    }
  }
}

// min/max over the element and its up to 8 neighbours (neighbor_minmax, viscosity_mod.F90:748-816), one thread per plane scalar
__global__ void __launch_bounds__(256) k_nbr_minmax(Geo G, int Q, const double* __restrict__ lmin, const double* __restrict__ lmax,
                                                    const double* __restrict__ ghost_mm, double* __restrict__ qmin,
                                                    double* __restrict__ qmax) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)G.ngroups * NKC * Q * GPL;
  if (p >= total) return;
  const int kk = p % KC, el = (p / KC) % GE;
  const size_t r = p / GPL;
  const int q = r % Q;
  const size_t gk = r / Q;
  const int kc = gk % NKC, g = gk / NKC;
  const int e = g * GE + el, k = kc * KC + kk;
  if (e >= G.nelem) return;
  double mn = lmin[p], mx = lmax[p];
  const int* nb = G.nbr8 + (size_t)e * 8;
  TSE_UNROLL
  for (int d = 0; d < 8; ++d) {
    const int b = nb[d];
    if (b >= 0) {
      const size_t pb = qplane(b, q, k, Q);
      mn = dmin(mn, lmin[pb]);
      mx = dmax(mx, lmax[pb]);
    } else if (b <= -2) {
      const size_t gb = ((size_t)(-b - 2) * 2 * Q + q) * NLEV + k;
      mn = dmin(mn, ghost_mm[gb]);
      mx = dmax(mx, ghost_mm[gb + (size_t)Q * NLEV]);
    }
  }
  qmin[p] = mn;
  qmax[p] = mx;
}

}  // namespace tse
