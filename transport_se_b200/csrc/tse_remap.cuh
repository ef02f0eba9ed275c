// vertical_remap + remap_Q_ppm (reference src/share/prim_advection_mod.F90:1242-1330, 98-356), PPM with
// mirrored boundary cells (vert_remap_q_alg != 2).
//
// One CTA per element; a thread owns one (column, tracer) pair and walks the 72 levels of that column serially, the way the
// reference's loops do.  Lanes 0..15 of a warp are the 16 columns of one tracer, lanes 16..31 those of the next one, so
//   * every global access of a warp is two full 128-byte lines (16 nodes x 8 B of one (level, tracer) plane each),
//   * everything that does not depend on the tracer (reciprocal source thickness, the PPM grid coefficients of
//     compute_ppm_grids, the search result kid / integration weights of every target interface) is computed once per element
//     into shared memory as [level][column] and read back as broadcasts: the lanes of the two tracers of a warp read the same
//     16 addresses, one wavefront per 64-bit and two per 128-bit load, no bank conflicts by construction,
//   * the PPM stencils (ao(j-2..j), dma(j-2..j-1), ai(j-3..j-2)) live in registers as sliding windows, the cumulative mass is a
//     running sum in the reference's order, and the parabola of cell j-2 is consumed as soon as it exists by the target cells
//     whose lower interface lies in it (kid is monotone, so the emission is a merge of two sorted sequences); nothing that
//     depends on the tracer touches shared memory.
// The walk reads a ring of 4 levels ahead of the level it emits, and a target cell k is always emitted after source level k+1 has
// been read (kid(k) >= k-1 by the search rule :159-165), so the remap can run in place.
// 18 warps per SM (35 tracers x 16 columns) against 8 for the level-parallel version it replaces (which needed 181 KB for
// per-tracer gather arrays): 40.4 -> see profiles/ for the current number.
#pragma once
#include "tse_ops.cuh"

namespace tse {

constexpr int RM_MAX_THREADS = 576;  // 36 tracer slots x 16 columns (35 tracers in one pass); more tracers loop
constexpr int RM_PF = 4;             // levels read ahead of the PPM stencil (one level chunk)
// shared memory, in doubles; every array is [index][16 columns]
constexpr int RS_DPO = 0;                         // [76]: dpo(j), j = -1..74 at index j+1
constexpr int RS_RDPO = RS_DPO + 76 * 16;         // [72]: 1/dpo(j), j = 1..72 at j-1
constexpr int RS_PX = RS_RDPO + 72 * 16;          // [75][4] double2, row j = what step j of the walk needs: (px0,px1)(j-1), (px2(j-1), px3(j-2)),
                                                  // (px4, px5*(px6-px7))(j-2), (px8,px9)(j-2); rows 1..74
constexpr int RS_PIO = RS_PX + 75 * 4 * 2 * 16;   // [74]: pio(1..74) at index 0..73 (last = sentinel)
constexpr int RS_PIN = RS_PIO + 74 * 16;          // [73]: pin(1..73)
constexpr int RS_G = RS_PIN + 73 * 16;            // [72][4]: g1, g2, g3 (integration weights of target interface k+1), dpo(kid(k))
constexpr int RS_END = RS_G + 72 * 4 * 16;
constexpr size_t RM_SMEM = (size_t)RS_END * sizeof(double) + 80 * 16 + 74 * 16 * sizeof(int);  // + kid[72][16] (bytes, unused rows pad to 80), cnt[74][16] (int)
static_assert(NLEV == 72 && KC == 4 && RM_PF == KC, "the level walk is unrolled by level chunks");

struct RemapArgs {
  double* q;                 // tracer field of time level np1_qdp (resolved), remapped in place
  const double* dp;          // derived%dp
  const double* divdp_proj;  // derived%divdp_proj
  double* dp3d;              // out: state%dp3d(:,:,:,np1)   [level field]
  double* ps_v;              // out: state%ps_v(:,:,np1)     [e][16]
  double dA[NLEV];           // (hyai(k+1)-hyai(k))*ps0   (by value: read from the constant bank inside a serial sum)
  double dB[NLEV];           // hybi(k+1)-hybi(k)
  double hyai0_ps0;          // hyai(1)*ps0
  double dt;
  int Q;
  int nelem;
  int* error_flag;           // set to 1 on negative layer thickness (prim_advection_mod.F90:1323)
};

// stage 1 of compute_ppm (:284-295): limited slope of cell j from ao(j-1), ao(j), ao(j+1) and (px0, px1, px2) of level j
__device__ __forceinline__ double ppm_dma(double am, double a0, double ap, double2 p01, double p2) {
  const double dl = a0 - am, dr = ap - a0;
  const double da = p01.x * (p01.y * dr + p2 * dl);
  double d = dmin(fabs(da), dmin(2. * fabs(dl), 2. * fabs(dr)));
  d = copysign(d, da);
  if (dr * dl <= 0.) d = 0.;
  return d;
}

__global__ void __launch_bounds__(RM_MAX_THREADS, 1) k_vertical_remap(const __grid_constant__ RemapArgs a) {
  extern __shared__ double sm[];
  const int e = blockIdx.x;
  if (e >= a.nelem) return;
  const int t = threadIdx.x, nthr = blockDim.x;
  const int n = t & 15;  // column (node of the 4x4 plane)
  double* const s_dpo = sm + RS_DPO;
  double* const s_rdpo = sm + RS_RDPO;
  double2* const s_px = reinterpret_cast<double2*>(sm + RS_PX);
  double* const s_pio = sm + RS_PIO;
  double* const s_pin = sm + RS_PIN;
  double* const s_g = sm + RS_G;
  unsigned char* const s_kid = reinterpret_cast<unsigned char*>(sm + RS_END);  // [72][16]: kid(k) (1-based cell index)
  int* const s_cnt = reinterpret_cast<int*>(s_kid + 80 * 16);                  // [74][16]: number of target cells whose kid == j

  // ---- grids: everything below is tracer independent -----------------------------------------------------------------
  bool neg = false;
  for (int i = t; i < NLEV * 16; i += nthr) {
    const int k = i >> 4;  // level k+1, column n (i & 15 == n because nthr is a multiple of 16)
    const size_t lp = lplane(e, k) * 16 + n;
    const double d = a.dp[lp] - a.dt * a.divdp_proj[lp];  // dp3d(np1) = dp_star (:1310-1313)
    a.dp3d[lp] = d;
    s_dpo[(k + 2) * 16 + n] = d;
    // mirrored ghost cells (:147-150): dpo(0) = dpo(1), dpo(-1) = dpo(2), dpo(nlev+1) = dpo(nlev), dpo(nlev+2) = dpo(nlev-1)
    if (k == 0) s_dpo[1 * 16 + n] = d;
    if (k == 1) s_dpo[0 * 16 + n] = d;
    if (k == NLEV - 1) s_dpo[74 * 16 + n] = d;
    if (k == NLEV - 2) s_dpo[75 * 16 + n] = d;
    neg |= (d < 0.0);
  }
  for (int i = t; i < 74 * 16; i += nthr) s_cnt[i] = 0;
  // the reference aborts here (prim_advection_mod.F90:1319-1324); the grid search below needs monotone pressures
  if (__syncthreads_or(neg)) {
    if (t == 0) *a.error_flag = 1;
    return;
  }
  if (t < 32) {
    // Sequential sums in the reference's order, one lane per column: lanes 0..15 ps_v and pin, lanes 16..31 pio.  The other
    // warps compute the PPM grid coefficients meanwhile (they depend on dpo only).
    if (t < 16) {
      double s = 0.0;
#pragma unroll 8
      for (int k = 1; k <= NLEV; ++k) s += s_dpo[(k + 1) * 16 + n];  // sum(dp3d,3)
      const double ps = a.hyai0_ps0 + s;
      a.ps_v[(size_t)e * 16 + n] = ps;
      double pin = 0.0;
      s_pin[0 * 16 + n] = 0.0;
#pragma unroll 8
      for (int k = 1; k < NLEV; ++k) {
        pin += a.dA[k - 1] + a.dB[k - 1] * ps;  // dp = dA*ps0 + dB*ps_v (:1314-1316)
        s_pin[k * 16 + n] = pin;                // pin(k+1) at index k
      }
      s_pin[NLEV * 16 + n] = s;  // pin(nlev+1) = pio(nlev+1) (:144): the same sequential sum as pio
    } else {
      double pio = 0.0;
      s_pio[0 * 16 + n] = 0.0;
#pragma unroll 8
      for (int k = 1; k <= NLEV; ++k) {
        pio += s_dpo[(k + 1) * 16 + n];
        s_pio[k * 16 + n] = pio;  // pio(k+1) at index k
      }
      s_pio[(NLEV + 1) * 16 + n] = pio + 1.0;  // sentinel (:141)
    }
  }
  // compute_ppm_grids (:221-260) for every (level j = 0..73, column); dx(j) = s_dpo[j+1].  Warp 0 joins after its sums when the
  // CTA is a single warp.
  {
    const int t0 = nthr > 32 ? t - 32 : t, nt = nthr > 32 ? nthr - 32 : nthr;
    for (int i = t0; i >= 0 && i < 74 * 16; i += nt) {
      const int j = i >> 4;
      const double dm = s_dpo[j * 16 + n], d0 = s_dpo[(j + 1) * 16 + n], d1 = s_dpo[(j + 2) * 16 + n];
      // stage-1 coefficients of level j are used by step j+1 (dma(j)), stage-2 coefficients by step j+2 (ai(j))
      double* r1 = reinterpret_cast<double*>(s_px + (size_t)(j + 1) * 4 * 16 + n);
      r1[0] = d0 / (dm + d0 + d1);              // pair 0
      r1[1] = (2. * dm + d0) / (d1 + d0);
      r1[32] = (d0 + 2. * d1) / (dm + d0);      // pair 1, first half
      if (j >= 1 && j <= NLEV) s_rdpo[(j - 1) * 16 + n] = 1.0 / d0;
      if (j <= NLEV) {
        const double d2 = s_dpo[(j + 3) * 16 + n];
        double* r2 = reinterpret_cast<double*>(s_px + (size_t)(j + 2) * 4 * 16 + n);
        const double r01 = 1.0 / (d0 + d1), r201 = 1.0 / (2. * d0 + d1), r021 = 1.0 / (d0 + 2. * d1);
        r2[33] = d0 * r01;                        // px3                     pair 1, second half
        r2[64] = 1. / (dm + d0 + d1 + d2);        // px4                     pair 2
        // the reference's dx(6)*(dx(7)-dx(8)) (:303), evaluated once.  (Shared reciprocals instead of 6 more divisions: the
        // coefficients differ from the reference's by <= 2 ulp.)
        r2[65] = (2. * d1 * d0) * r01 * ((dm + d0) * r201 - (d2 + d1) * r021);
        r2[96] = d0 * (dm + d0) * r201;           // px8                     pair 3
        r2[97] = d1 * (d1 + d2) * r021;           // px9
      }
    }
  }
  __syncthreads();
  // search (:155-173): kid(k) = old cell holding new interface k+1, z2 its normalised position; integration weights of
  // integrate_parabola (:349-356) with x1 = -0.5
  for (int i = t; i < NLEV * 16; i += nthr) {
    const int k = (i >> 4) + 1;
    int kk = k;
    const double pk = s_pin[k * 16 + n];                 // pin(k+1)
    while (s_pio[(kk - 1) * 16 + n] <= pk) ++kk;         // pio(kk)
    --kk;
    if (kk == NLEV + 1) kk = NLEV;
    const double dk = s_dpo[(kk + 1) * 16 + n];
    const double x2 = (pk - (s_pio[(kk - 1) * 16 + n] + s_pio[kk * 16 + n]) * 0.5) / dk, x1 = -0.5;
    atomicAdd(s_cnt + kk * 16 + n, 1);  // target cells per source cell (kid is monotone: the emission below is a merge)
    double* g = s_g + (size_t)(k - 1) * 4 * 16 + n;
    g[0] = x2 - x1;
    g[16] = (x2 * x2 - x1 * x1) * 0.5;
    g[32] = (x2 * x2 * x2 - x1 * x1 * x1) / 3.0;
    g[48] = dk;
  }
  __syncthreads();

  // ---- tracers ---------------------------------------------------------------------------------------------------------
  const double2* const px = s_px + n;
  const int skc = a.Q * GPL * 16;  // doubles between consecutive level chunks of one (element, tracer)
  for (int q = t >> 4; q < a.Q; q += nthr >> 4) {
    double* const col = a.q + qplane(e, q, 0, a.Q) * 16 + n;  // (level 1, node n) of this tracer
    // start-up: ao(1), ao(2), the mirrored cells ao(0) = ao(1), ao(-1) = ao(2) (:193-196), dma(0), dma(1), ai(0)
    double r1 = col[0], r2 = col[16];                           // raw Qdp of cells j-2, j-1 (for the cumulative mass)
    double ring[RM_PF] = {col[32], col[48], col[skc], col[skc + 16]};  // levels 3..6, read ahead of the stencil
    double a1 = r1 * s_rdpo[n], a2 = r2 * s_rdpo[16 + n];       // window: a1 = ao(j-2), a2 = ao(j-1) at the top of step j
    double dma_prev, ai_prev;
    {
      const double2 q01 = px[1 * 64], q2 = px[1 * 64 + 16];     // stage 1 of level 0
      const double dma0 = ppm_dma(a2, a1, a1, q01, q2.x);
      const double2 s01 = px[2 * 64], s23 = px[2 * 64 + 16], s45 = px[2 * 64 + 32], s89 = px[2 * 64 + 48];  // stage 1 of level 1, stage 2 of level 0
      dma_prev = ppm_dma(a1, a1, a2, s01, s23.x);               // dma(1)
      ai_prev = a1 + s23.y * (a1 - a1) + s45.x * (s45.y * (a1 - a1) - s89.x * dma_prev + s89.y * dma0);  // ai(0) (:300-305)
    }
    double masso = 0.0;   // masso(j-2): mass above cell j-2
    double massn1 = 0.0;
    int k = 0;            // next target cell (0-based)

    // One step of the walk: newest source value a3 = ao(j), coefficient row = step j.  Produces the parabola of cell j-2
    // (c0, c1, c2); straight-line code, so that the four steps of a round interleave in the instruction stream.
    auto step = [&](double raw, double a3, const double2* row, double& c0, double& c1, double& c2, double& mo) {
      const double2 p01 = row[0], p23 = row[16], p45 = row[32], p89 = row[48];
      // dma(j-1) from ao(j-2), ao(j-1), ao(j) and (px0,px1,px2) of level j-1
      const double dma = ppm_dma(a1, a2, a3, p01, p23.x);
      // ai(j-2) (:300-305) with the stage-2 coefficients of level j-2
      const double ai = a1 + p23.y * (a2 - a1) + p45.x * (p45.y * (a2 - a1) - p89.x * dma + p89.y * dma_prev);
      // parabola of cell j-2 (:310-333): aj = ao(j-2), al = ai(j-3), ar = ai(j-2)
      const double aj = a1;
      double al = ai_prev, ar = ai;
      if ((ar - aj) * (aj - al) <= 0.) {
        al = aj;
        ar = aj;
      }
      // the reference divides by 6 (:323-329); multiplying by the rounded reciprocal differs by <= 1 ulp
      if ((ar - al) * (aj - (al + ar) * 0.5) > (ar - al) * (ar - al) * (1.0 / 6.0)) al = 3. * aj - 2. * ar;
      if ((ar - al) * (aj - (al + ar) * 0.5) < -((ar - al) * (ar - al)) * (1.0 / 6.0)) ar = 3. * aj - 2. * al;
      c0 = 1.5 * aj - (al + ar) * 0.25;
      c1 = ar - al;
      c2 = -6. * aj + 3. * (al + ar);
      mo = masso;  // masso(j-2)
      masso += r1;
      r1 = r2;
      r2 = raw;
      a1 = a2;
      a2 = a3;
      dma_prev = dma;
      ai_prev = ai;
    };
    // target cells whose lower interface lies in the cell of this parabola: massn2 = masso(kid) + integral*dpo(kid) (:201-209)
    auto emit = [&](int cnt, double c0, double c1, double c2, double mo) {
#pragma unroll 1
      for (; cnt > 0; --cnt) {
        const double* g = s_g + k * 64 + n;
        const double integ = c0 * g[0] + c1 * g[16] + c2 * g[32];
        const double massn2 = mo + integ * g[48];
        col[(k >> 2) * skc + (k & 3) * 16] = massn2 - massn1;
        massn1 = massn2;
        ++k;
      }
    };

    // steps j = 3 + 4 i + u; source level j sits in chunk (2 + 4 i + u) / 4 at position (2 + u) % 4
    const double* src = col + skc;  // chunk 1 + i: the ring is refilled from levels j + 4
#pragma unroll 1
    for (int i = 0; i < 17; ++i) {
      const int j = 3 + 4 * i;
      const double2* row = px + j * 64;
      const double* rd = s_rdpo + (j - 1) * 16 + n;
      const int* cn = s_cnt + (j - 2) * 16 + n;
      const double w0 = ring[0], w1 = ring[1], w2 = ring[2], w3 = ring[3];
      ring[0] = src[32];
      ring[1] = src[48];
      if (i < 16) {
        ring[2] = src[skc];
        ring[3] = src[skc + 16];
      }
      double c[4][3], mo[4];
      step(w0, w0 * rd[0], row, c[0][0], c[0][1], c[0][2], mo[0]);
      step(w1, w1 * rd[16], row + 64, c[1][0], c[1][1], c[1][2], mo[1]);
      step(w2, w2 * rd[32], row + 128, c[2][0], c[2][1], c[2][2], mo[2]);
      step(w3, w3 * rd[48], row + 192, c[3][0], c[3][1], c[3][2], mo[3]);
      TSE_UNROLL
      for (int u = 0; u < 4; ++u) emit(cn[16 * u], c[u][0], c[u][1], c[u][2], mo[u]);
      src += skc;
    }
    {  // last round, j = 71..74: two real levels, then the mirrored cells ao(73) = ao(72), ao(74) = ao(71)
      const double2* row = px + 71 * 64;
      const double* rd = s_rdpo + 70 * 16 + n;
      const int* cn = s_cnt + 69 * 16 + n;
      double c[4][3], mo[4];
      step(ring[0], ring[0] * rd[0], row, c[0][0], c[0][1], c[0][2], mo[0]);
      step(ring[1], ring[1] * rd[16], row + 64, c[1][0], c[1][1], c[1][2], mo[1]);
      const double a71 = a1;                    // at the top of step 73: a1 = ao(71), a2 = ao(72)
      step(0.0, a2, row + 128, c[2][0], c[2][1], c[2][2], mo[2]);
      step(0.0, a71, row + 192, c[3][0], c[3][1], c[3][2], mo[3]);
      TSE_UNROLL
      for (int u = 0; u < 4; ++u) emit(cn[16 * u], c[u][0], c[u][1], c[u][2], mo[u]);
    }
  }
}

}  // namespace tse
