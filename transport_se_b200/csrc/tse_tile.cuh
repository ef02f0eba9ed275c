// Tiled tracer-field kernels (sm_100a): the production versions of the plane-per-thread kernels in tse_kernels.cuh.
//
// CTA = (group of 16 elements, chunk of 4 levels); it walks all tracers, QI=2 at a time, through a double-buffered
// cp.async pipeline.  For each step the 16 KB tile [q][el][kk][16] (contiguous in HBM) is copied with coalesced 16-byte
// cp.async into shared memory, XOR-swizzled per 128-byte plane so that a thread can read "its" plane with conflict-free
// 128-bit loads.  One thread owns one plane: all 4x4 contractions, the limiter and the extrema are register-only.
//
// DSS (edgeVpack / bndry_exchangeV / edgeVunpack, edge_mod.F90:366-742) is fused into the load of the consumer: a field is
// stored "pre-DSS" (spheremp-weighted); neighbours inside the group are read straight from the tile in shared memory,
// neighbours outside the group (the patch perimeter, ~68 nodes for a 4x4 patch) are fetched by 8-byte cp.async into a halo
// array in the same pipeline stage, and the sum runs in the reference's unpack order.  The rspheremp factor of the DSS is
// folded into the per-level package (E1, U, rdp below), computed once per CTA and reused for all tracers.
//
// Per-(element, level) package (registers/shared memory, tracer independent), with rX = rspheremp if the input is pre-DSS else 1:
//   dp_s = dp - rhs_mult*dt*divdp_proj, Vstar = vn0/dp_s, dp_star = dp_s - dt*divdp        (prim_advection_mod.F90:753,847-864)
//   U_c  = rX * metdet*(Dinv(c,1)*Vstar1 + Dinv(c,2)*Vstar2)      gv_c = U_c * S       (S = raw DSS sum)
//   E1   = spheremp*rX, E2 = dt*spheremp*rmetdet*rrearth          y = spheremp*Qtens = E1*S - E2*div
//   CL   = spheremp*dp_star  (the limiter's c), RDP = rX/dp_s     (Q = S*RDP for the extrema)
// The limiter works on y = c*x directly (limiter_y), so neither 1/dp_star nor the final spheremp multiply is needed.
#pragma once
#include "tse_kernels.cuh"

namespace tse {

constexpr int QI = 2;                  // tracers per pipeline step
constexpr int TT = QI * GPL;           // 128 threads
constexpr int TILE_BYTES = TT * 128;   // 16 KB

enum TileOp { OP_MINMAX = 0, OP_STAGE1, OP_STAGE2, OP_STAGE3, OP_BIHARM_PRE, OP_TIME_AVG, OP_RESOLVE };

struct TileTables {
  const int* gsrc_t;    // [npad][NSLOT]: <0 none, [0,256) in-group (el<<4|node), >=256 halo entry (code-256)
  const int* halo_off;  // [ngroups+1]
  const int* halo_src;  // [halo_off[ngroups]]: >=0 (elem<<4|node), <=-2 ghost slot
  int hmax;             // max halo entries of a group
};

struct TileArgs {
  const double* src[2];    // input fields: [0] main (Qdp), [1] second (STAGE3: qtens is src[0], Qdp is src[1]; TIME_AVG: q0 is src[0])
  int pending[2];
  const double* ghost[2];
  double* out;
  const double *vn0, *dp, *divdp, *divdp_proj;
  double rhs_mult_dt, dt, visc_coef, rkstage;
  const double* dp0;
  double *qmin, *qmax, *qmin_loc, *qmax_loc;
  int Q;
};

__host__ __device__ inline int tile_stage_bytes(int hmax) { return TILE_BYTES + QI * hmax * KC * 8 + 16; }
constexpr int PP_BYTES = GPL * 128;  // per-plane package field (8 KB)
constexpr int EL_BYTES = GE * 128;   // per-element package field (2 KB)
enum { PP_U1 = 0, PP_U2, PP_CL, PP_RDP, NPP };
enum { EL_E1 = 0, EL_E2, EL_RSPH, EL_T11, EL_T12, EL_T22, NEL };
__host__ __device__ inline int tile_smem_bytes(int hmax) { return 2 * tile_stage_bytes(hmax) + NPP * PP_BYTES + NEL * EL_BYTES; }

__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ double2 lds128(const unsigned char* base, int off) { return *reinterpret_cast<const double2*>(base + off); }
__device__ __forceinline__ double lds64(const unsigned char* base, int off) { return *reinterpret_cast<const double*>(base + off); }

// per-plane package field: 16-byte unit index c*64 + pl
__device__ __forceinline__ void ld_pp(const unsigned char* f, int pl, double (&v)[16]) {
  TSE_UNROLL
  for (int c = 0; c < 8; ++c) {
    const double2 t = lds128(f, (c * GPL + pl) * 16);
    v[2 * c] = t.x;
    v[2 * c + 1] = t.y;
  }
}
// per-element package field: 16-byte unit index c*16 + el
__device__ __forceinline__ void ld_el(const unsigned char* f, int el, double (&v)[16]) {
  TSE_UNROLL
  for (int c = 0; c < 8; ++c) {
    const double2 t = lds128(f, (c * GE + el) * 16);
    v[2 * c] = t.x;
    v[2 * c + 1] = t.y;
  }
}

// limiter_optim_iter_full (prim_advection_mod.F90:976-1094) on y = c*x (mass contributions) instead of x:
// x > maxp  <=>  y > maxp*c ; addmass += (x-maxp)*c = y - maxp*c ; x += inc  <=>  y += inc*c ; result ptens*sphweights = y.
__device__ __forceinline__ void limiter_y(double (&y)[16], const double (&c)[16], double sumc, double& minp, double& maxp) {
  const double tol_limiter = (double)5e-14f;
  if (sumc <= 0.0) return;
  double mass = 0.0;
  TSE_UNROLL
  for (int k1 = 0; k1 < 16; ++k1) mass += y[(k1 >> 2) + 4 * (k1 & 3)];
  if (mass < minp * sumc) minp = mass / sumc;
  if (mass > maxp * sumc) maxp = mass / sumc;
  const double thresh = tol_limiter * fabs(mass);
  for (int iter = 1; iter <= NPSQ - 1; ++iter) {
    double addmass = 0.0;
    TSE_UNROLL
    for (int k1 = 0; k1 < 16; ++k1) {
      const int n = (k1 >> 2) + 4 * (k1 & 3);
      const double hi = maxp * c[n], lo = minp * c[n];
      if (y[n] > hi) {
        addmass += y[n] - hi;
        y[n] = hi;
      }
      if (y[n] < lo) {
        addmass -= lo - y[n];
        y[n] = lo;
      }
    }
    if (fabs(addmass) <= thresh) break;
    double weightssum = 0.0;
    if (addmass > 0.0) {
      TSE_UNROLL
      for (int k1 = 0; k1 < 16; ++k1) {
        const int n = (k1 >> 2) + 4 * (k1 & 3);
        if (y[n] < maxp * c[n]) weightssum += c[n];
      }
      const double inc = addmass / weightssum;
      TSE_UNROLL
      for (int n = 0; n < 16; ++n)
        if (y[n] < maxp * c[n]) y[n] = fma(inc, c[n], y[n]);
    } else {
      TSE_UNROLL
      for (int k1 = 0; k1 < 16; ++k1) {
        const int n = (k1 >> 2) + 4 * (k1 & 3);
        if (y[n] > minp * c[n]) weightssum += c[n];
      }
      const double inc = addmass / weightssum;
      TSE_UNROLL
      for (int n = 0; n < 16; ++n)
        if (y[n] > minp * c[n]) y[n] = fma(inc, c[n], y[n]);
    }
  }
}

// laplace_sphere_wk with the per-element tensor T read from the element-level package
__device__ __forceinline__ void laplace_wk_el(const double (&s)[16], const Dvv& D, const unsigned char* el_base, int el, double (&lap)[16]) {
  double w1[16], w2[16];
  {
    double d1[16], d2[16];
    grad_raw(s, D, d1, d2);
    TSE_UNROLL
    for (int c = 0; c < 8; ++c) {
      const double2 a = lds128(el_base + EL_T11 * EL_BYTES, (c * GE + el) * 16);
      const double2 b = lds128(el_base + EL_T12 * EL_BYTES, (c * GE + el) * 16);
      const double2 cc = lds128(el_base + EL_T22 * EL_BYTES, (c * GE + el) * 16);
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
__global__ void __launch_bounds__(TT, 2) k_tile(Geo G, Dvv D, TileTables tb, TileArgs a) {
  constexpr bool kStage = (OP == OP_STAGE1 || OP == OP_STAGE2 || OP == OP_STAGE3);
  constexpr int NIN = (OP == OP_STAGE3 || OP == OP_TIME_AVG) ? 2 : 1;
  constexpr bool kHasOut = (OP != OP_MINMAX);
  extern __shared__ __align__(16) unsigned char smem[];
  const int SB = tile_stage_bytes(tb.hmax);
  unsigned char* const pp = smem + 2 * SB;
  unsigned char* const elb = pp + NPP * PP_BYTES;
  const int ZERO_OFF = TILE_BYTES + QI * tb.hmax * KC * 8;

  const int t = threadIdx.x;
  const int g = blockIdx.x / NKC, kc = blockIdx.x % NKC;
  const int qi = t / GPL, pl = t % GPL, el = pl / KC, kk = pl % KC;
  const int e = g * GE + el, k = kc * KC + kk;
  const bool evalid = e < G.nelem;
  const int Q = a.Q;
  const int hoff = tb.halo_off[g], H = tb.halo_off[g + 1] - hoff;

  // ---- level package -------------------------------------------------------------------------------------------
  const bool main_pending = (OP == OP_STAGE3) ? (a.pending[1] != 0) : (a.pending[0] != 0);
  if (t < 2) *reinterpret_cast<double*>(smem + t * SB + ZERO_OFF) = 0.0;
  {
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
      const double* T = G.T + (size_t)ee * 48 + 2 * c;
      t11 = *reinterpret_cast<const double2*>(T);
      t12 = *reinterpret_cast<const double2*>(T + 16);
      t22 = *reinterpret_cast<const double2*>(T + 32);
    }
    const int off = (c * GE + pe) * 16;
    *reinterpret_cast<double2*>(elb + EL_E1 * EL_BYTES + off) = e1;
    *reinterpret_cast<double2*>(elb + EL_E2 * EL_BYTES + off) = e2;
    *reinterpret_cast<double2*>(elb + EL_RSPH * EL_BYTES + off) = rs;
    *reinterpret_cast<double2*>(elb + EL_T11 * EL_BYTES + off) = t11;
    *reinterpret_cast<double2*>(elb + EL_T12 * EL_BYTES + off) = t12;
    *reinterpret_cast<double2*>(elb + EL_T22 * EL_BYTES + off) = t22;
  }
  if (kStage || OP == OP_MINMAX || OP == OP_BIHARM_PRE) {
    // per-plane fields: thread -> (plane t>>1, nodes 8*(t&1) .. +7)
    const int ppl = t >> 1, half = t & 1;
    const int pe = g * GE + ppl / KC, pk = kc * KC + ppl % KC;
    TSE_UNROLL
    for (int cc = 0; cc < 4; ++cc) {
      const int c = half * 4 + cc, n = 2 * c;
      double2 u1 = make_double2(0, 0), u2 = u1, cl = make_double2(1, 1), rd = make_double2(1, 1);
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
        }
      }
      const int off = (c * GPL + ppl) * 16;
      *reinterpret_cast<double2*>(pp + PP_U1 * PP_BYTES + off) = u1;
      *reinterpret_cast<double2*>(pp + PP_U2 * PP_BYTES + off) = u2;
      *reinterpret_cast<double2*>(pp + PP_CL * PP_BYTES + off) = cl;
      *reinterpret_cast<double2*>(pp + PP_RDP * PP_BYTES + off) = rd;
    }
  }

  // ---- per-thread DSS gather offsets (bytes inside a stage buffer) ------------------------------------------------
  int goff[NSLOT];
  {
    const int* gs = tb.gsrc_t + (size_t)(evalid ? e : 0) * NSLOT;
    TSE_UNROLL
    for (int s = 0; s < NSLOT; ++s) {
      const int code = evalid ? gs[s] : -1;
      int off = ZERO_OFF;
      if (code >= 256) off = TILE_BYTES + ((qi * tb.hmax + (code - 256)) * KC + kk) * 8;
      else if (code >= 0) {
        const int p2 = qi * GPL + (code >> 4) * KC + kk, node = code & 15;
        off = p2 * 128 + ((((node >> 1) ^ (p2 & 7))) << 4) + (node & 1) * 8;
      }
      goff[s] = off;
    }
  }
  const unsigned smem_u32 = (unsigned)__cvta_generic_to_shared(smem);

  const int nit = (Q + QI - 1) / QI;
  const int nitems = nit * NIN;
  auto issue = [&](int j) {
    const int it = j / NIN, which = j % NIN;
    const double* src = a.src[which];
    const int q0 = it * QI, nq = min(QI, Q - q0);
    const unsigned dst = smem_u32 + (j & 1) * SB;
    const size_t base = (((size_t)g * NKC + kc) * Q + q0) * GPL * 16;
    TSE_UNROLL
    for (int r = 0; r < 8; ++r) {
      const int i = r * TT + t, p = i >> 3, c = i & 7;
      if (p < nq * GPL) cp_async16(dst + p * 128 + ((c ^ (p & 7)) << 4), src + base + (size_t)i * 2);
    }
    if (a.pending[which]) {
      const int total = nq * KC * H;
      for (int idx = t; idx < total; idx += TT) {
        const int h = idx % H, r2 = idx / H, kk2 = r2 % KC, qi2 = r2 / KC;
        const int code = tb.halo_src[hoff + h];
        const int kq = kc * KC + kk2, qq = q0 + qi2;
        const double* gp = (code >= 0) ? src + (qplane(code >> 4, qq, kq, Q) * 16 + (code & 15))
                                       : a.ghost[which] + ((size_t)(-code - 2) * Q + qq) * NLEV + kq;
        cp_async8(dst + TILE_BYTES + ((qi2 * tb.hmax + h) * KC + kk2) * 8, gp);
      }
    }
    cp_async_commit();
  };

  // q-invariant per-thread values
  double sumc = 0.0;
  __syncthreads();  // package visible
  if (kStage) {
    TSE_UNROLL
    for (int k1 = 0; k1 < 16; ++k1) {
      const int n = (k1 >> 2) + 4 * (k1 & 3);
      sumc += lds64(pp + PP_CL * PP_BYTES, ((n >> 1) * GPL + pl) * 16 + (n & 1) * 8);
    }
  }
  const double cf = (OP == OP_STAGE3) ? a.visc_coef * a.dp0[k] : 0.0;

  double keep[16];  // STAGE3: cf*lap of the first item; TIME_AVG: Qdp(n0)
  issue(0);
  for (int j = 0; j < nitems; ++j) {
    if (j + 1 < nitems) {
      issue(j + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int it = j / NIN, which = j % NIN;
    const int q = it * QI + qi;
    const bool valid = evalid && q < Q;
    const unsigned char* buf = smem + (j & 1) * SB;
    const size_t pidx = (((size_t)g * NKC + kc) * Q + q) * GPL + pl;  // global plane index

    double S[16];
    TSE_UNROLL
    for (int c = 0; c < 8; ++c) {
      const double2 v = lds128(buf, t * 128 + ((c ^ (t & 7)) << 4));
      S[2 * c] = v.x;
      S[2 * c + 1] = v.y;
    }
    if (a.pending[which]) {  // DSS in the reference's unpack order: S, E, N, W edges, then SW, SE, NE, NW corners
      TSE_UNROLL
      for (int i = 0; i < 4; ++i) S[i] += lds64(buf, goff[i]);
      TSE_UNROLL
      for (int i = 0; i < 4; ++i) S[3 + 4 * i] += lds64(buf, goff[4 + i]);
      TSE_UNROLL
      for (int i = 0; i < 4; ++i) S[12 + i] += lds64(buf, goff[8 + i]);
      TSE_UNROLL
      for (int i = 0; i < 4; ++i) S[4 * i] += lds64(buf, goff[12 + i]);
      S[0] += lds64(buf, goff[16]);
      S[3] += lds64(buf, goff[17]);
      S[15] += lds64(buf, goff[18]);
      S[12] += lds64(buf, goff[19]);
    }
    const bool last_of_iter = (which == NIN - 1);
    if (kHasOut && last_of_iter) __syncthreads();  // everyone has read this buffer; it becomes the output staging tile

    if (valid) {
      if (OP == OP_MINMAX || OP == OP_BIHARM_PRE) {
        double mn, mx;
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) {
          const double2 rd = lds128(pp + PP_RDP * PP_BYTES, (c * GPL + pl) * 16);
          S[2 * c] *= rd.x;
          S[2 * c + 1] *= rd.y;
        }
        mn = S[0];
        mx = S[0];
        TSE_UNROLL
        for (int n = 1; n < 16; ++n) {
          mn = fmin(mn, S[n]);
          mx = fmax(mx, S[n]);
        }
        a.qmin_loc[pidx] = mn;
        a.qmax_loc[pidx] = mx;
        if (OP == OP_BIHARM_PRE) {
          double lap[16];
          laplace_wk_el(S, D, elb, el, lap);
          TSE_UNROLL
          for (int n = 0; n < 16; ++n) S[n] = lap[n];
        }
      } else if (OP == OP_RESOLVE) {
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) {
          const double2 rs = lds128(elb + EL_RSPH * EL_BYTES, (c * GE + el) * 16);
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
            double2 rs = lds128(elb + EL_RSPH * EL_BYTES, (c * GE + el) * 16);
            if (!a.pending[1]) rs = make_double2(1.0, 1.0);
            S[2 * c] = (keep[2 * c] + (a.rkstage - 1.0) * (rs.x * S[2 * c])) / a.rkstage;
            S[2 * c + 1] = (keep[2 * c + 1] + (a.rkstage - 1.0) * (rs.y * S[2 * c + 1])) / a.rkstage;
          }
        }
      } else if (OP == OP_STAGE3 && which == 0) {
        // second half of biharmonic_wk_scalar_minmax: lap(rspheremp*DSS(qtens)); Qtens_biharmonic*spheremp = cf*lap
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) {
          const double2 rs = lds128(elb + EL_RSPH * EL_BYTES, (c * GE + el) * 16);
          S[2 * c] *= rs.x;
          S[2 * c + 1] *= rs.y;
        }
        double lap[16];
        laplace_wk_el(S, D, elb, el, lap);
        TSE_UNROLL
        for (int n = 0; n < 16; ++n) keep[n] = cf * lap[n];
      } else if (kStage) {
        double minp = a.qmin[pidx], maxp = a.qmax[pidx];
        if (OP == OP_STAGE2) {
          double mn, mx;
          {
            const double2 rd = lds128(pp + PP_RDP * PP_BYTES, pl * 16);
            mn = S[0] * rd.x;
            mx = mn;
            const double q1 = S[1] * rd.y;
            mn = fmin(mn, q1);
            mx = fmax(mx, q1);
          }
          TSE_UNROLL
          for (int c = 1; c < 8; ++c) {
            const double2 rd = lds128(pp + PP_RDP * PP_BYTES, (c * GPL + pl) * 16);
            const double q0v = S[2 * c] * rd.x, q1v = S[2 * c + 1] * rd.y;
            mn = fmin(mn, fmin(q0v, q1v));
            mx = fmax(mx, fmax(q0v, q1v));
          }
          minp = fmin(minp, mn);
          maxp = fmax(maxp, mx);
        }
        double y[16];
        {
          double g1[16], g2[16];
          TSE_UNROLL
          for (int c = 0; c < 8; ++c) {
            const double2 u1 = lds128(pp + PP_U1 * PP_BYTES, (c * GPL + pl) * 16);
            const double2 u2 = lds128(pp + PP_U2 * PP_BYTES, (c * GPL + pl) * 16);
            g1[2 * c] = u1.x * S[2 * c];
            g1[2 * c + 1] = u1.y * S[2 * c + 1];
            g2[2 * c] = u2.x * S[2 * c];
            g2[2 * c + 1] = u2.y * S[2 * c + 1];
          }
          div_contract(g1, g2, D, y);
        }
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) {
          const double2 e1 = lds128(elb + EL_E1 * EL_BYTES, (c * GE + el) * 16);
          const double2 e2 = lds128(elb + EL_E2 * EL_BYTES, (c * GE + el) * 16);
          y[2 * c] = fma(-e2.x, y[2 * c], e1.x * S[2 * c]);
          y[2 * c + 1] = fma(-e2.y, y[2 * c + 1], e1.y * S[2 * c + 1]);
          if (OP == OP_STAGE3) {
            y[2 * c] += keep[2 * c];
            y[2 * c + 1] += keep[2 * c + 1];
          }
        }
        {
          double c[16];
          ld_pp(pp + PP_CL * PP_BYTES, pl, c);
          limiter_y(y, c, sumc, minp, maxp);
        }
        a.qmin[pidx] = minp;
        a.qmax[pidx] = maxp;
        TSE_UNROLL
        for (int n = 0; n < 16; ++n) S[n] = y[n];
      }
    }

    if (kHasOut && last_of_iter) {
      if (valid) {
        unsigned char* wb = smem + (j & 1) * SB;
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) *reinterpret_cast<double2*>(wb + t * 128 + ((c ^ (t & 7)) << 4)) = make_double2(S[2 * c], S[2 * c + 1]);
      }
      __syncthreads();
      const int q0 = it * QI, nq = min(QI, Q - q0);
      const size_t base = (((size_t)g * NKC + kc) * Q + q0) * GPL * 16;
      TSE_UNROLL
      for (int r = 0; r < 8; ++r) {
        const int i = r * TT + t, p = i >> 3, c = i & 7;
        if (p < nq * GPL && g * GE + ((p / KC) % GE) < G.nelem)
          *reinterpret_cast<double2*>(a.out + base + (size_t)i * 2) = lds128(buf, p * 128 + ((c ^ (p & 7)) << 4));
      }
    }
    __syncthreads();  // buffer (j&1) is free for the load issued at the top of the next iteration
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
      mn = fmin(mn, lmin[pb]);
      mx = fmax(mx, lmax[pb]);
    } else if (b <= -2) {
      const size_t gb = ((size_t)(-b - 2) * 2 * Q + q) * NLEV + k;
      mn = fmin(mn, ghost_mm[gb]);
      mx = fmax(mx, ghost_mm[gb + (size_t)Q * NLEV]);
    }
  }
  qmin[p] = mn;
  qmax[p] = mx;
}

}  // namespace tse
