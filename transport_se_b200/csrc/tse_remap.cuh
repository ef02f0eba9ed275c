// vertical_remap + remap_Q_ppm (reference src/share/prim_advection_mod.F90:1242-1330, 98-356), PPM with
// mirrored boundary cells (vert_remap_q_alg != 2).
//
// One CTA per element, 16 columns x 32 level-lanes.  Everything that does not depend on the tracer
// (source/target grids, the search kid/z2, the 10 PPM grid coefficients) is computed once per element and
// kept in registers of the thread that owns the level; the tracer loop then runs level-parallel through
// shared memory (the PPM stencils are local in k).  The only serial pieces are the prefix sums.
#pragma once
#include "tse_ops.cuh"

namespace tse {

constexpr int RM_KL = 32;                 // level lanes
constexpr int RM_THREADS = 16 * RM_KL;    // 512
constexpr int RM_R = 3;                   // levels owned per thread: j = kl + 32 r, j in 0..73
constexpr int RM_ROWS = 76 + 73 + 76 + 74 + 73 + 3 * 72 + 73 + 73 + 8;
constexpr int RM_SEG = 9;                 // the prefix sum over 72 levels runs as 8 segments of 9 levels
constexpr size_t RM_SMEM = (size_t)RM_ROWS * 16 * sizeof(double);

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
  double(*s_dpo)[16] = reinterpret_cast<double(*)[16]>(sm);  // row j+1, j=-1..74
  double(*s_araw)[16] = s_dpo + 76;                          // row k, k=0..72
  double(*s_ao)[16] = s_araw + 73;                           // row j+1
  double(*s_dma)[16] = s_ao + 76;                            // row j, j=0..73   (phase A: pio[0..73])
  double(*s_ai)[16] = s_dma + 74;                            // row j, j=0..72   (phase A: pin[0..72])
  double(*s_coef)[16] = s_ai + 73;                           // row c*72 + (j-1)
  double(*s_masso)[16] = s_coef + 216;                       // row k, k=0..72
  double(*s_m2)[16] = s_masso + 73;                          // row k, k=0..72: massn2
  double(*s_seg)[16] = s_m2 + 73;                            // row s, s=0..7: segment totals of the prefix sum
  double(*s_pio)[16] = s_dma;
  double(*s_pin)[16] = s_ai;

  const int e = blockIdx.x;
  if (e >= a.nelem) return;
  const int n = threadIdx.x & 15, kl = threadIdx.x >> 4;

  // ---- phase A: grids -------------------------------------------------------------------
  bool neg = false;
  for (int k = 1 + kl; k <= NLEV; k += RM_KL) {
    const size_t lp = lplane(e, k - 1) * 16 + n;
    const double d = a.dp[lp] - a.dt * a.divdp_proj[lp];  // dp3d(np1) = dp_star (:1310-1313)
    a.dp3d[lp] = d;
    s_dpo[k + 1][n] = d;
    neg |= (d < 0.0);
  }
  // the reference aborts here (prim_advection_mod.F90:1319-1324); the grid search below needs monotone pressures
  if (__syncthreads_or(neg)) {
    if (threadIdx.x == 0) *a.error_flag = 1;
    return;
  }
  if (kl == 0) {
    double s = 0.0;
    for (int k = 1; k <= NLEV; ++k) s += s_dpo[k + 1][n];  // sum(dp3d,3)
    const double ps = a.hyai0_ps0 + s;
    a.ps_v[(size_t)e * 16 + n] = ps;
    double pin = 0.0, pio = 0.0;
    s_pin[0][n] = 0.0;
    s_pio[0][n] = 0.0;
    for (int k = 1; k <= NLEV; ++k) {
      pin += a.dA[k - 1] + a.dB[k - 1] * ps;  // dp = dA*ps0 + dB*ps_v
      pio += s_dpo[k + 1][n];
      s_pin[k][n] = pin;
      s_pio[k][n] = pio;
    }
    s_pio[NLEV + 1][n] = pio + 1.0;  // sentinel (:147)
    s_pin[NLEV][n] = pio;            // pin(nlev+1) = pio(nlev+1) (:144)
    // mirrored ghost cells (:147-150)
    s_dpo[0][n] = s_dpo[3][n];   // dpo(-1) = dpo(2)
    s_dpo[1][n] = s_dpo[2][n];   // dpo(0)  = dpo(1)
    s_dpo[74][n] = s_dpo[73][n];  // dpo(nlev+1) = dpo(nlev)
    s_dpo[75][n] = s_dpo[72][n];  // dpo(nlev+2) = dpo(nlev-1)
  }
  __syncthreads();

  int kid[RM_R];
  double z2[RM_R], rdpo[RM_R], px[RM_R][10];
#pragma unroll
  for (int r = 0; r < RM_R; ++r) {
    const int j = kl + RM_KL * r;
    kid[r] = 1;
    z2[r] = 0.0;
    rdpo[r] = 0.0;
    if (j >= 1 && j <= NLEV) {
      int kk = j;
      const double pk = s_pin[j][n];
      while (s_pio[kk - 1][n] <= pk) ++kk;
      --kk;
      if (kk == NLEV + 1) kk = NLEV;
      kid[r] = kk;
      z2[r] = (pk - (s_pio[kk - 1][n] + s_pio[kk][n]) * 0.5) / s_dpo[kk + 1][n];
      rdpo[r] = 1.0 / s_dpo[j + 1][n];
    }
    if (j <= NLEV + 1) {  // compute_ppm_grids (:221-260); dx(j) = s_dpo[j+1]
      const double dm = s_dpo[j][n], d0 = s_dpo[j + 1][n], d1 = s_dpo[j + 2][n];
      px[r][0] = d0 / (dm + d0 + d1);
      px[r][1] = (2. * dm + d0) / (d1 + d0);
      px[r][2] = (d0 + 2. * d1) / (dm + d0);
      if (j <= NLEV) {
        const double d2 = s_dpo[j + 3][n];
        px[r][3] = d0 / (d0 + d1);
        px[r][4] = 1. / (dm + d0 + d1 + d2);
        px[r][5] = (2. * d1 * d0) / (d0 + d1);
        px[r][6] = (dm + d0) / (2. * d0 + d1);
        px[r][7] = (d2 + d1) / (2. * d1 + d0);
        px[r][8] = d0 * (dm + d0) / (2. * d0 + d1);
        px[r][9] = d1 * (d1 + d2) / (d0 + 2. * d1);
      }
    }
  }
  __syncthreads();  // pio/pin storage is reused as dma/ai below

  // ---- phase B: tracers -----------------------------------------------------------------
  // The next tracer's column is prefetched into registers while the current one is processed.
  double nxt[RM_R];
  size_t goff[RM_R];   // offset of (e, tracer 0, level j-1, node n); consecutive tracers are GPL*16 doubles apart
  bool own[RM_R];      // this thread owns level j = kl + 32 r in 1..72
#pragma unroll
  for (int r = 0; r < RM_R; ++r) {
    const int j = kl + RM_KL * r;
    own[r] = (j >= 1 && j <= NLEV);
    goff[r] = own[r] ? qplane(e, 0, j - 1, a.Q) * 16 + n : 0;
    nxt[r] = own[r] ? a.q[goff[r]] : 0.0;
  }
  const double third = 1.0 / 3.0, sixth = 1.0 / 6.0;
  const int sg = threadIdx.x >> 4;  // prefix-sum segment handled by threads 0..127 (column n, segment sg)
  for (int q = 0; q < a.Q; ++q) {
#pragma unroll
    for (int r = 0; r < RM_R; ++r) {
      const int j = kl + RM_KL * r;
      if (own[r]) {
        s_araw[j][n] = nxt[r];
        s_ao[j + 1][n] = nxt[r] * rdpo[r];  // ao = Qdp/dpo (:187)
      }
    }
    if (q + 1 < a.Q) {
#pragma unroll
      for (int r = 0; r < RM_R; ++r)
        if (own[r]) nxt[r] = a.q[goff[r] + (size_t)(q + 1) * (GPL * 16)];
    }
    __syncthreads();
    // masso prefix sum (:184-186) as 8 segments of 9 levels: segment totals, then offsets (fixed order, deterministic)
    if (sg < 8) {
      double t = 0.0;
#pragma unroll
      for (int i = 1; i <= RM_SEG; ++i) t += s_araw[sg * RM_SEG + i][n];
      s_seg[sg][n] = t;
    } else if (sg == 8) {  // mirrored ghost cells (:193-196)
      s_ao[0][n] = s_ao[3][n];
      s_ao[1][n] = s_ao[2][n];
      s_ao[74][n] = s_ao[73][n];
      s_ao[75][n] = s_ao[72][n];
    }
    __syncthreads();
    if (sg < 8) {
      double m = 0.0;
      for (int i = 0; i < sg; ++i) m += s_seg[i][n];
      if (sg == 0) s_masso[0][n] = 0.0;
#pragma unroll
      for (int i = 1; i <= RM_SEG; ++i) {
        m += s_araw[sg * RM_SEG + i][n];
        s_masso[sg * RM_SEG + i][n] = m;
      }
    }
    // compute_ppm (:267-342): dma
#pragma unroll
    for (int r = 0; r < RM_R; ++r) {
      const int j = kl + RM_KL * r;
      if (j <= NLEV + 1) {
        const double am = s_ao[j][n], a0 = s_ao[j + 1][n], ap = s_ao[j + 2][n];
        const double da = px[r][0] * (px[r][1] * (ap - a0) + px[r][2] * (a0 - am));
        double d = dmin(fabs(da), dmin(2. * fabs(a0 - am), 2. * fabs(ap - a0)));
        d = copysign(d, da);
        if ((ap - a0) * (a0 - am) <= 0.) d = 0.;
        s_dma[j][n] = d;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RM_R; ++r) {
      const int j = kl + RM_KL * r;
      if (j <= NLEV) {
        const double a0 = s_ao[j + 1][n], ap = s_ao[j + 2][n];
        s_ai[j][n] = a0 + px[r][3] * (ap - a0) +
                     px[r][4] * (px[r][5] * (px[r][6] - px[r][7]) * (ap - a0) - px[r][8] * s_dma[j + 1][n] + px[r][9] * s_dma[j][n]);
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RM_R; ++r) {
      const int j = kl + RM_KL * r;
      if (own[r]) {
        const double aj = s_ao[j + 1][n];
        double al = s_ai[j - 1][n], ar = s_ai[j][n];
        if ((ar - aj) * (aj - al) <= 0.) {
          al = aj;
          ar = aj;
        }
        // the reference divides by 6 (:323-329); multiplying by the rounded reciprocal differs by <= 1 ulp
        if ((ar - al) * (aj - (al + ar) * 0.5) > (ar - al) * (ar - al) * sixth) al = 3. * aj - 2. * ar;
        if ((ar - al) * (aj - (al + ar) * 0.5) < -((ar - al) * (ar - al)) * sixth) ar = 3. * aj - 2. * al;
        s_coef[j - 1][n] = 1.5 * aj - (al + ar) * 0.25;
        s_coef[72 + j - 1][n] = ar - al;
        s_coef[144 + j - 1][n] = -6. * aj + 3. * (al + ar);
      }
    }
    __syncthreads();
    // massn2(k) = masso(kid) + integral over the part of cell kid below the new interface (:201-209)
#pragma unroll
    for (int r = 0; r < RM_R; ++r) {
      const int j = kl + RM_KL * r;
      if (own[r]) {
        const int kk = kid[r];
        const double x1 = -0.5, x2 = z2[r];
        const double c0 = s_coef[kk - 1][n], c1 = s_coef[72 + kk - 1][n], c2 = s_coef[144 + kk - 1][n];
        const double integ = c0 * (x2 - x1) + c1 * (x2 * x2 - x1 * x1) * 0.5 + c2 * (x2 * x2 * x2 - x1 * x1 * x1) * third;
        s_m2[j][n] = s_masso[kk - 1][n] + integ * s_dpo[kk + 1][n];
      } else if (j == 0) {
        s_m2[0][n] = 0.0;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RM_R; ++r) {
      const int j = kl + RM_KL * r;
      if (own[r]) a.q[goff[r] + (size_t)q * (GPL * 16)] = s_m2[j][n] - s_m2[j - 1][n];
    }
    // no barrier needed here: the next iteration first writes s_araw/s_ao, whose last readers sit before the previous barrier
  }
}

}  // namespace tse
