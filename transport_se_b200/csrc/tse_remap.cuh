// vertical_remap + remap_Q_ppm (reference src/share/prim_advection_mod.F90:1242-1330, 98-356), PPM with
// mirrored boundary cells (vert_remap_q_alg != 2).
//
// One CTA per element, 256 threads = 2 tracers in flight x 128 threads.  A thread owns a run of 9 consecutive levels of one
// column (72 = 8 segments x 9): the PPM stencils (ao(j-2..j+2), dma(j), dma(j+1), ai(j-1), ai(j)) then live in registers as
// sliding windows, the only neighbour traffic is the two ghost values on either side of the run (shuffles between the 8 lanes
// that hold a column) and the segment carries of the cumulative mass.  Everything that does not depend on the tracer is set up
// once per element: source thickness and its reciprocal, the search result (kid, z2) of each target interface (registers), and the
// PPM grid coefficients of compute_ppm_grids (8 doubles per (column, level), shared memory, read as 4 x 128-bit per level).
// Shared memory is otherwise only used for what is gathered with a data-dependent index: the parabola coefficients and the
// cumulative mass of cell kid(k).  Columns are warp-private (a warp holds 4 columns x 8 segments), so the tracer loop needs no
// CTA barrier; HBM is read and written straight from registers (4 consecutive nodes x 8 levels = 8 sectors per warp request),
// with the next tracer's run prefetched.
#pragma once
#include "tse_ops.cuh"

namespace tse {

constexpr int RM_NSEG = 8;                 // segments per column
constexpr int RM_L = NLEV / RM_NSEG;       // 9 levels per segment
#ifndef TSE_RM_TQ
#define TSE_RM_TQ 2
#endif
constexpr int RM_TQ = TSE_RM_TQ;           // tracers in flight per CTA
constexpr int RM_THREADS = RM_TQ * 16 * RM_NSEG;  // 256
static_assert(RM_L * RM_NSEG == NLEV, "levels split evenly into segments");
// shared memory (doubles)
// Bank layout: the 8 lanes of a column sit 9 levels apart.  For 64-bit accesses that is conflict-free when the column stride is
// 8 mod 16 doubles (72, 88); the 128-bit px reads get a per-(column, segment) copy of their 11 levels, 90 doubles apart.
constexpr int RM_PXS = 11 * 8 + 2;                // px block of one (column, segment): levels j0..j0+10, 8 doubles each, + pad
constexpr int RM_LD = 88;                         // column stride of dpo / masso
constexpr int RS_PX = 0;                          // [16][8][11][8]: px0, px1, px2, px3, px4, px5*(px6-px7), px8, px9
constexpr int RS_DPO = RS_PX + 16 * RM_NSEG * RM_PXS;  // [16][RM_LD]: dpo(j), j = -1..74 at index j+1
constexpr int RS_TR = RS_DPO + 16 * RM_LD;        // per tracer in flight: c0,c1,c2 [16][72] (cell j at j-1), masso [16][RM_LD]
constexpr int RS_TR_SIZE = 16 * (3 * 72 + RM_LD);
constexpr size_t RM_SMEM = (size_t)(RS_TR + RM_TQ * RS_TR_SIZE) * sizeof(double);

struct RemapArgs {
  double* q;                 // tracer field of time level np1_qdp (resolved), remapped in place
  const double* dp;          // derived%dp
  const double* divdp_proj;  // derived%divdp_proj
  double* dp3d;              // out: state%dp3d(:,:,:,np1)   [level field]
  double* ps_v;              // out: state%ps_v(:,:,np1)     [e][16]
  const double* dA;          // [NLEV] (hyai(k+1)-hyai(k))*ps0
  const double* dB;          // [NLEV] hybi(k+1)-hybi(k)
  double hyai0_ps0;          // hyai(1)*ps0
  double dt;
  int Q;
  int nelem;
  int* error_flag;           // set to 1 on negative layer thickness (prim_advection_mod.F90:1323)
};

__global__ void __launch_bounds__(RM_THREADS, 1) k_vertical_remap(RemapArgs a) {
  extern __shared__ double sm[];
  const int e = blockIdx.x;
  if (e >= a.nelem) return;
  const int t = threadIdx.x;

  // ---- phase A: grids (all 256 threads) ------------------------------------------------------------------------------
  double* const s_dpo = sm + RS_DPO;
  // pio/pin live in the (not yet used) per-tracer area during the set-up
  double* const s_pio = sm + RS_TR;             // [16][75]: pio(0..73) (73 = sentinel)
  double* const s_pin = sm + RS_TR + 16 * 75;   // [16][73]: pin(0..72)
  bool neg = false;
  {
    const int ln = t & 15, lk = t >> 4;  // coalesced mapping: 16 nodes x 16 level lanes
    for (int k = 1 + lk; k <= NLEV; k += 16) {
      const size_t lp = lplane(e, k - 1) * 16 + ln;
      const double d = a.dp[lp] - a.dt * a.divdp_proj[lp];  // dp3d(np1) = dp_star (:1310-1313)
      a.dp3d[lp] = d;
      s_dpo[ln * RM_LD + k + 1] = d;
      neg |= (d < 0.0);
    }
  }
  // the reference aborts here (prim_advection_mod.F90:1319-1324); the grid search below needs monotone pressures
  if (__syncthreads_or(neg)) {
    if (t == 0) *a.error_flag = 1;
    return;
  }
  if (t < 16) {  // sequential sums in the reference's order, one thread per column (once per element)
    const int n = t;
    double* dpo = s_dpo + n * RM_LD;
    double s = 0.0;
    for (int k = 1; k <= NLEV; ++k) s += dpo[k + 1];  // sum(dp3d,3)
    const double ps = a.hyai0_ps0 + s;
    a.ps_v[(size_t)e * 16 + n] = ps;
    double pin = 0.0, pio = 0.0;
    s_pin[n * 73] = 0.0;
    s_pio[n * 75] = 0.0;
    for (int k = 1; k <= NLEV; ++k) {
      pin += a.dA[k - 1] + a.dB[k - 1] * ps;  // dp = dA*ps0 + dB*ps_v
      pio += dpo[k + 1];
      s_pin[n * 73 + k] = pin;
      s_pio[n * 75 + k] = pio;
    }
    s_pio[n * 75 + NLEV + 1] = pio + 1.0;  // sentinel (:147)
    s_pin[n * 73 + NLEV] = pio;            // pin(nlev+1) = pio(nlev+1) (:144)
    // mirrored ghost cells (:147-150)
    dpo[0] = dpo[3];    // dpo(-1) = dpo(2)
    dpo[1] = dpo[2];    // dpo(0)  = dpo(1)
    dpo[74] = dpo[73];  // dpo(nlev+1) = dpo(nlev)
    dpo[75] = dpo[72];  // dpo(nlev+2) = dpo(nlev-1)
  }
  __syncthreads();
  // compute_ppm_grids (:221-260) for every (column, level j = 0..73); dx(j) = dpo[j+1]
  for (int i = t; i < 16 * RM_NSEG * 11; i += RM_THREADS) {
    const int n = i / (RM_NSEG * 11), sg = (i / 11) % RM_NSEG, li = i % 11;
    const int j = RM_L * sg + li;  // levels j0..j0+10 of segment sg (the two runs next to a boundary both hold its levels)
    const double* dpo = s_dpo + n * RM_LD;
    const double dm = dpo[j], d0 = dpo[j + 1], d1 = dpo[j + 2];
    double px[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    px[0] = d0 / (dm + d0 + d1);
    px[1] = (2. * dm + d0) / (d1 + d0);
    px[2] = (d0 + 2. * d1) / (dm + d0);
    if (j <= NLEV) {
      const double d2 = dpo[j + 3];
      px[3] = d0 / (d0 + d1);
      px[4] = 1. / (dm + d0 + d1 + d2);
      const double p5 = (2. * d1 * d0) / (d0 + d1), p6 = (dm + d0) / (2. * d0 + d1), p7 = (d2 + d1) / (2. * d1 + d0);
      px[5] = p5 * (p6 - p7);  // the reference's dx(5)*(dx(6)-dx(7)) (:303), evaluated once
      px[6] = d0 * (dm + d0) / (2. * d0 + d1);
      px[7] = d1 * (d1 + d2) / (d0 + 2. * d1);
    }
    double2* dst = reinterpret_cast<double2*>(sm + RS_PX + (n * RM_NSEG + sg) * RM_PXS + li * 8);
    dst[0] = make_double2(px[0], px[1]);
    dst[1] = make_double2(px[2], px[3]);
    dst[2] = make_double2(px[4], px[5]);
    dst[3] = make_double2(px[6], px[7]);
  }

  // ---- per-thread, tracer-independent: the run of levels j = j0+1 .. j0+9 of column n ---------------------------------
  const int half = t / 128;                 // which of the RM_TQ tracers in flight
  const int tw = t % 128, wv = tw >> 5, lane = tw & 31;
  const int seg = lane & 7, n = 4 * wv + (lane >> 3);
  const int j0 = RM_L * seg;
  const double* dpo = s_dpo + n * RM_LD;
  int kid[RM_L];
  double z2[RM_L], rdpo[RM_L];
  TSE_UNROLL
  for (int i = 0; i < RM_L; ++i) {
    const int j = j0 + 1 + i;
    int kk = j;
    const double pk = s_pin[n * 73 + j];
    while (s_pio[n * 75 + kk - 1] <= pk) ++kk;
    --kk;
    if (kk == NLEV + 1) kk = NLEV;
    kid[i] = kk;
    z2[i] = (pk - (s_pio[n * 75 + kk - 1] + s_pio[n * 75 + kk]) * 0.5) / dpo[kk + 1];
    rdpo[i] = 1.0 / dpo[j + 1];
  }
  __syncthreads();  // pio/pin are dead: their storage becomes the per-tracer arrays; px is complete

  // ---- phase B: tracers -----------------------------------------------------------------------------------------------
  double* const s_c0 = sm + RS_TR + half * RS_TR_SIZE + n * 72;
  double* const s_c1 = s_c0 + 16 * 72;
  double* const s_c2 = s_c1 + 16 * 72;
  double* const s_masso = sm + RS_TR + half * RS_TR_SIZE + 3 * 16 * 72 + n * RM_LD;
  const double2* const s_px = reinterpret_cast<const double2*>(sm + RS_PX + (n * RM_NSEG + seg) * RM_PXS);  // level j0 first
  const unsigned FULL = 0xffffffffu;  // shuffles run with width 8: the 8 lanes that hold this column
  size_t goff[RM_L];  // (e, tracer 0, level j-1, node n); consecutive tracers are GPL*16 doubles apart
  TSE_UNROLL
  for (int i = 0; i < RM_L; ++i) goff[i] = qplane(e, 0, j0 + i, a.Q) * 16 + n;
  const double third = 1.0 / 3.0, sixth = 1.0 / 6.0;
  double nxt[RM_L];
  TSE_UNROLL
  for (int i = 0; i < RM_L; ++i) nxt[i] = (half < a.Q) ? a.q[goff[i] + (size_t)half * (GPL * 16)] : 0.0;
  // both halves run the same number of iterations (warps are not split between halves, but this keeps the loop simple): the
  // last iteration of the second half is a dry run when Q is odd
  const int niter = (a.Q + RM_TQ - 1) / RM_TQ;
  for (int it = 0; it < niter; ++it) {
    const int q = it * RM_TQ + half;
    const bool live = q < a.Q;
    // ao = Qdp/dpo (:187), window w[i+2] = ao(j0+1+i), i = -2..10
    double w[RM_L + 4], raw[RM_L];
    TSE_UNROLL
    for (int i = 0; i < RM_L; ++i) {
      raw[i] = nxt[i];
      w[i + 2] = raw[i] * rdpo[i];
    }
    if (q + RM_TQ < a.Q) {
      TSE_UNROLL
      for (int i = 0; i < RM_L; ++i) nxt[i] = a.q[goff[i] + (size_t)(q + RM_TQ) * (GPL * 16)];
    }
    // cumulative mass masso (:184-186): running sum inside the run, segment carries added in segment order (fixed, deterministic)
    {
      double run[RM_L];
      double m = 0.0;
      TSE_UNROLL
      for (int i = 0; i < RM_L; ++i) {
        m += raw[i];
        run[i] = m;
      }
      double base = 0.0;
      TSE_UNROLL
      for (int s2 = 0; s2 < RM_NSEG - 1; ++s2) {
        const double tot = __shfl_sync(FULL, m, s2, 8);
        if (s2 < seg) base += tot;
      }
      if (seg == 0) s_masso[0] = 0.0;
      TSE_UNROLL
      for (int i = 0; i < RM_L; ++i) s_masso[j0 + 1 + i] = base + run[i];
    }
    // ghost values from the neighbouring runs; mirrored cells at the column ends (:193-196)
    {
      const double up1 = __shfl_up_sync(FULL, w[RM_L + 1], 1, 8), up2 = __shfl_up_sync(FULL, w[RM_L], 1, 8);
      const double dn1 = __shfl_down_sync(FULL, w[2], 1, 8), dn2 = __shfl_down_sync(FULL, w[3], 1, 8);
      w[1] = seg == 0 ? w[2] : up1;                               // ao(j0)    | ao(0)  = ao(1)
      w[0] = seg == 0 ? w[3] : up2;                               // ao(j0-1)  | ao(-1) = ao(2)
      w[RM_L + 2] = seg == RM_NSEG - 1 ? w[RM_L + 1] : dn1;       // ao(j0+10) | ao(73) = ao(72)
      w[RM_L + 3] = seg == RM_NSEG - 1 ? w[RM_L] : dn2;           // ao(j0+11) | ao(74) = ao(71)
    }
    // compute_ppm (:267-342) level by level: dma(j), then ai(j-1) (needs dma(j-1), dma(j)), then the parabola of cell j-1
    double d_prev = 0.0, ai_prev = 0.0;
    double2 pxa_prev = make_double2(0, 0), pxb_prev = pxa_prev, pxc_prev = pxa_prev;
    TSE_UNROLL
    for (int i = -1; i <= RM_L; ++i) {  // level j = j0 + 1 + i (j0 .. j0+10); the window index of ao(j) is i + 2
      const double2 p01 = s_px[(i + 1) * 4], p23 = s_px[(i + 1) * 4 + 1];
      const double am = w[i + 1], a0 = w[i + 2], ap = w[i + 3];
      const double da = p01.x * (p01.y * (ap - a0) + p23.x * (a0 - am));
      double d = dmin(fabs(da), dmin(2. * fabs(a0 - am), 2. * fabs(ap - a0)));
      d = copysign(d, da);
      if ((ap - a0) * (a0 - am) <= 0.) d = 0.;
      if (i >= 0) {
        // ai(j-1) with the coefficients of level j-1 (kept from the previous pass); its a0 = ao(j-1) = am, its ap = ao(j) = a0
        const double ai = am + pxa_prev.y * (a0 - am) + pxb_prev.x * (pxb_prev.y * (a0 - am) - pxc_prev.x * d + pxc_prev.y * d_prev);
        if (i >= 1) {
          // parabola of cell j-1 = j0 + i (:310-333): aj = ao(j-1) = am, al = ai(j-2), ar = ai(j-1)
          const double aj = am;
          double al = ai_prev, ar = ai;
          if ((ar - aj) * (aj - al) <= 0.) {
            al = aj;
            ar = aj;
          }
          // the reference divides by 6 (:323-329); multiplying by the rounded reciprocal differs by <= 1 ulp
          if ((ar - al) * (aj - (al + ar) * 0.5) > (ar - al) * (ar - al) * sixth) al = 3. * aj - 2. * ar;
          if ((ar - al) * (aj - (al + ar) * 0.5) < -((ar - al) * (ar - al)) * sixth) ar = 3. * aj - 2. * al;
          s_c0[j0 + i - 1] = 1.5 * aj - (al + ar) * 0.25;
          s_c1[j0 + i - 1] = ar - al;
          s_c2[j0 + i - 1] = -6. * aj + 3. * (al + ar);
        }
        ai_prev = ai;
      }
      d_prev = d;
      pxa_prev = p23;                 // (px2, px3)
      pxb_prev = s_px[(i + 1) * 4 + 2];  // (px4, px5*(px6-px7))
      pxc_prev = s_px[(i + 1) * 4 + 3];  // (px8, px9)
    }
    __syncwarp();
    // massn2(k) = masso(kid) + integral over the part of cell kid below the new interface (:201-209)
    double m2[RM_L];
    TSE_UNROLL
    for (int i = 0; i < RM_L; ++i) {
      const int kk = kid[i];
      const double x1 = -0.5, x2 = z2[i];
      const double c0 = s_c0[kk - 1], c1 = s_c1[kk - 1], c2 = s_c2[kk - 1];
      const double integ = c0 * (x2 - x1) + c1 * (x2 * x2 - x1 * x1) * 0.5 + c2 * (x2 * x2 * x2 - x1 * x1 * x1) * third;
      m2[i] = s_masso[kk - 1] + integ * dpo[kk + 1];
    }
    double below = __shfl_up_sync(FULL, m2[RM_L - 1], 1, 8);  // massn2(j0) from the run below
    if (seg == 0) below = 0.0;
    if (live) {
      TSE_UNROLL
      for (int i = 0; i < RM_L; ++i) a.q[goff[i] + (size_t)q * (GPL * 16)] = m2[i] - (i == 0 ? below : m2[i - 1]);
    }
    __syncwarp();  // the next tracer overwrites the column's coefficient / masso arrays
  }
}

}  // namespace tse
