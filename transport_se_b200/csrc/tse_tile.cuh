// Building blocks of the tiled tracer-field kernels (sm_100a); the kernel itself (k_pipe) is in tse_pipe.cuh.
//
// CTA = (group of 16 elements, chunk of 4 levels); it walks all tracers, QI at a time.  For each step the QI*8 KB tile
// [q][el][kk][16] (contiguous in HBM) lands in shared memory XOR-swizzled per 128-byte plane, so that a thread can read "its"
// plane with conflict-free 128-bit loads.  One thread owns one plane: all 4x4 contractions, the limiter and the extrema are
// register-only.  A warp holds 8/QI elements x 4 levels x QI tracers, so that the data-dependent limiter loop diverges as
// little as possible (the behaviour of a plane is mostly a property of its element).
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
//   CL   = spheremp*dp_star  (the limiter's c), RC = 1/CL, RDP = rX/dp_s     (Q = S*RDP for the extrema)
// The limiter works on y = c*x (limiter_y), so neither 1/dp_star nor the final spheremp multiply is needed on its fast path.
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
#ifndef TSE_QW
#define TSE_QW (TSE_QI < 8 ? ((TSE_QI % 2) ? TSE_QI : 2) : 8)
#endif
constexpr int QW = TSE_QW;             // tracers per warp (2: a warp is 4 elements x 4 levels x 2 tracers)
constexpr int EPW = 8 / QW;            // elements per warp (1 for QI >= 8: the limiter's work is a property of the element)
static_assert(KC == 4 && (QI % QW) == 0 && (8 % QW) == 0 && (GE % EPW) == 0 && TT % 32 == 0 && TT <= 512, "warp mapping");
static_assert(EPW > 1 || (TT / 8) % GPL == 0 || GPL % (TT / 8) == 0, "swizzle");

// XOR swizzle of the 16-byte units of a 128-byte plane: distinct over the 8 planes a quarter-warp reads together
// (lanes 0..7 = 4 levels x 2 elements for EPW > 1, 4 levels x 2 tracers for EPW == 1)
__host__ __device__ constexpr int swz(int p) { return EPW > 1 ? (p & 7) : ((p & 3) | (((p / GPL) & 1) << 2)); }

enum TileOp { OP_MINMAX = 0, OP_STAGE1, OP_STAGE2, OP_STAGE3, OP_BIHARM_PRE, OP_TIME_AVG, OP_RESOLVE, OP_MASS, OP_HYPERVIS };
// OP_HYPERVIS = OP_STAGE3's data flow (second laplacian of the first input, added to the second) with the zero limiter instead of
// limiter 8: its own instantiation so that the hot stage kernels do not carry the extra code (tse_advance_hypervis_scalar)

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
  int limiter8;      // 1: limiter_option == 8 (limiter_optim_iter_full); 0: no limiter inside euler_step, as the reference for every other value
  int store_bounds;  // stage ops: write the limiter's relaxed qmin/qmax back (needed after stage 1; otherwise only for inspection)
  int zero;          // always 0 (a value the compiler cannot fold: see mbar_arrive_after in tse_pipe.cuh)
  // OP_MASS (tse_diag_mass): fixed-point accumulators [2*Q], running max of |J| as bits [Q], binary shift per tracer [Q]
  long long* mass_acc;
  unsigned long long* mass_maxbits;
  const int* mass_shift;
  const int* glist;  // optional list of groups this launch covers (boundary groups first, interior groups while the halo is in flight)
};

// per-plane package field: 8 chunks (16 bytes = 2 nodes each) x GPL planes, chunk-major.  The chunk stride is padded by one
// 16-byte unit: with a stride of exactly GPL*16 = 1 KB the 8 chunks of a plane fall into the same banks and the prologue's stores
// (8 consecutive threads = the 8 chunks of one plane, so that the global loads are whole 128-byte lines) went 8-way conflicted
// (7.6 % of all shared-memory wavefronts of a stage kernel); the consumers' reads (consecutive planes of one chunk) are
// conflict-free either way.
constexpr int PCS = GPL * 16 + 16;   // bytes between consecutive chunks of a package field
constexpr int PP_BYTES = 8 * PCS;
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
       : (op == OP_STAGE3 || op == OP_HYPERVIS) ? TileCfg{4, 6, 1, 0, 1, 2, -1, 3, 0, 1, 2, 3, 4, 5}
       : op == OP_MINMAX ? TileCfg{1, 0, 0, -1, -1, -1, 0, -1, -1, -1, -1, -1, -1, -1}
       : op == OP_BIHARM_PRE ? TileCfg{1, 3, 1, -1, -1, -1, 0, -1, -1, -1, -1, 0, 1, 2}
       : op == OP_MASS ? TileCfg{0, 2, 0, -1, -1, -1, -1, -1, 0, -1, 1, -1, -1, -1}   // E1 slot = spheremp (unscaled), RSPH
                       : TileCfg{0, 1, 1, -1, -1, -1, -1, -1, -1, -1, 0, -1, -1, -1};
}
// one IN stage: tile, halo [tracer][halo node][level], a zero (target of absent DSS neighbours), limiter bounds qmin|qmax [tracer][plane]
constexpr int BND_BYTES = 2 * QI * GPL * 8;
__host__ __device__ constexpr int tile_in_bytes(int hmax) { return TILE_BYTES + QI * hmax * KC * 8 + 16 + BND_BYTES; }

__device__ __forceinline__ void cp_async8(unsigned dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(dst), "l"(src));
}

__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ double2 lds128(const unsigned char* base, int off) { return *reinterpret_cast<const double2*>(base + off); }
__device__ __forceinline__ double lds64(const unsigned char* base, int off) { return *reinterpret_cast<const double*>(base + off); }

// volatile shared-memory load: keeps the limiter's c out of registers (re-read from the package each pass)
__device__ __forceinline__ double2 lds128v(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// 1 if any of the 16 values lies outside [lo, hi].  Four predicates OR-accumulate the 32 compares (setp.<cmp>.or): 2 DSETP per
// node and no selects.
__device__ __forceinline__ unsigned any_outside(const double (&x)[16], double lo, double hi) {
  unsigned viol;
    asm("{\n"
        ".reg .pred p0, p1, p2, p3;\n"
        "setp.gt.f64 p0, %1, %17;\n    setp.lt.or.f64 p0, %1, %18, p0;\n"
        "setp.gt.f64 p1, %2, %17;\n    setp.lt.or.f64 p1, %2, %18, p1;\n"
        "setp.gt.f64 p2, %3, %17;\n    setp.lt.or.f64 p2, %3, %18, p2;\n"
        "setp.gt.f64 p3, %4, %17;\n    setp.lt.or.f64 p3, %4, %18, p3;\n"
        "setp.gt.or.f64 p0, %5, %17, p0;\n setp.lt.or.f64 p0, %5, %18, p0;\n"
        "setp.gt.or.f64 p1, %6, %17, p1;\n setp.lt.or.f64 p1, %6, %18, p1;\n"
        "setp.gt.or.f64 p2, %7, %17, p2;\n setp.lt.or.f64 p2, %7, %18, p2;\n"
        "setp.gt.or.f64 p3, %8, %17, p3;\n setp.lt.or.f64 p3, %8, %18, p3;\n"
        "setp.gt.or.f64 p0, %9, %17, p0;\n setp.lt.or.f64 p0, %9, %18, p0;\n"
        "setp.gt.or.f64 p1, %10, %17, p1;\n setp.lt.or.f64 p1, %10, %18, p1;\n"
        "setp.gt.or.f64 p2, %11, %17, p2;\n setp.lt.or.f64 p2, %11, %18, p2;\n"
        "setp.gt.or.f64 p3, %12, %17, p3;\n setp.lt.or.f64 p3, %12, %18, p3;\n"
        "setp.gt.or.f64 p0, %13, %17, p0;\n setp.lt.or.f64 p0, %13, %18, p0;\n"
        "setp.gt.or.f64 p1, %14, %17, p1;\n setp.lt.or.f64 p1, %14, %18, p1;\n"
        "setp.gt.or.f64 p2, %15, %17, p2;\n setp.lt.or.f64 p2, %15, %18, p2;\n"
        "setp.gt.or.f64 p3, %16, %17, p3;\n setp.lt.or.f64 p3, %16, %18, p3;\n"
        "or.pred p0, p0, p1;\n or.pred p2, p2, p3;\n or.pred p0, p0, p2;\n"
        "selp.u32 %0, 1, 0, p0;\n"
        "}\n"
        : "=r"(viol)
        : "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]), "d"(x[4]), "d"(x[5]), "d"(x[6]), "d"(x[7]), "d"(x[8]), "d"(x[9]), "d"(x[10]), "d"(x[11]),
          "d"(x[12]), "d"(x[13]), "d"(x[14]), "d"(x[15]), "d"(hi), "d"(lo));
  return viol;
}

// limiter_optim_iter_full (prim_advection_mod.F90:976-1094).  On entry y = c*x (mass contributions, c = sphweights*dpmass);
// c and rc = 1/c are read from the per-plane package in shared memory (cbase/rcbase = shared address of chunk 0 of this
// plane, chunk stride PCS bytes).  Sums use 4 interleaved partial accumulators (fixed order, identical on every GPU count).
//
// Fast path: mass = sum(y) and the min/max relaxation (:1016-1029) need no x; if no x = y*rc lies outside [minp, maxp] the
// reference's first sweep finds addmass = 0 and leaves (:1047), so y is returned untouched.
// Slow path (x in place of y): sweep 1 clips against both bounds.  From then on the direction is fixed: redistributing
// addmass > 0 raises nodes below maxp, so later sweeps can only find nodes above maxp and addmass stays >= 0 (and the mirror
// image for addmass < 0); the lower-bound test of the reference's sweeps 2..15 is then never taken.  Working on z = -x,
// bound -minp for the downward case (negation is exact) leaves one code path, whose sweep fuses "add the increment"
// (:1052-1078 of sweep i), "clip" (:1037-1045 of sweep i+1) and the next weightssum; the weight of a node is carried in a
// register and zeroed when the node reaches the bound, which removes the per-node "still below the bound?" tests.
// limiter2d_zero (cuda_mod.F90:863-913 / prim_advection_mod.F90:1186-1234) on one plane of spheremp-weighted values: flip the sign
// if the plane's mass is negative, zero the negative nodes, rescale the others to the original mass.  Sums in the reference's
// node order.
__device__ __forceinline__ void limiter2d_zero(double (&y)[16]) {
  double mass = 0.0;
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) mass = mass + y[n];
  const bool negm = mass < 0.0;
  double mass_new = 0.0;
  TSE_UNROLL
  for (int n = 0; n < 16; ++n) {
    double v = negm ? -y[n] : y[n];
    v = v < 0.0 ? 0.0 : v;
    mass_new = mass_new + v;
    y[n] = v;
  }
  if (mass_new > 0.0) {
    TSE_UNROLL
    for (int n = 0; n < 16; ++n) y[n] = y[n] * fabs(mass) / mass_new;
  }
  if (negm) {
    TSE_UNROLL
    for (int n = 0; n < 16; ++n) y[n] = -y[n];
  }
}

// limiter_check: mass, relaxation of the bounds, and the test "does any node leave [minp, maxp]".  Returns 1 if the slow path has
// to run (y is untouched either way).  Requires sumc > 0.
__device__ __forceinline__ unsigned limiter_check(const double (&y)[16], unsigned rcbase, double sumc, double& minp, double& maxp, double& mass) {
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
  // any x = y*rc outside [minp, maxp]?  (DMUL + 2 DSETP per node)
  double x[16];
  TSE_UNROLL
  for (int cc = 0; cc < 8; ++cc) {
    const double2 r = lds128v(rcbase + cc * PCS);
    x[2 * cc] = y[2 * cc] * r.x;
    x[2 * cc + 1] = y[2 * cc + 1] * r.y;
  }
  return any_outside(x, minp, maxp);
}

#ifdef TSE_EXP_LIMSTATS
__device__ unsigned long long g_lim_stats[4];
#endif
// limiter_slow: the clip / redistribute sweeps, for a plane limiter_check flagged (minp/maxp already relaxed, mass from the check).
__device__ __forceinline__ void limiter_slow(double (&y)[16], unsigned cbase, unsigned rcbase, double mass, double minp, double maxp) {
  const double tol_limiter = (double)5e-14f;
  // ---- slow path ---------------------------------------------------------------------------------------------------
  // x = y*rc in place of y.  Sweep 1 clips against both bounds (:1037-1045).
  const double thresh = tol_limiter * fabs(mass);
  TSE_UNROLL
  for (int cc = 0; cc < 8; ++cc) {
    const double2 r = lds128v(rcbase + cc * PCS);
    y[2 * cc] *= r.x;
    y[2 * cc + 1] *= r.y;
  }
  double ce[16];  // c of the nodes that can still be moved towards the bound (0 for the others); c itself until the direction is known
  double am;
  {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    TSE_UNROLL
    for (int cc = 0; cc < 8; cc += 2) {
      const double2 ca = lds128v(cbase + cc * PCS), cb = lds128v(cbase + (cc + 1) * PCS);
      const int n = 2 * cc;
      ce[n] = ca.x; ce[n + 1] = ca.y; ce[n + 2] = cb.x; ce[n + 3] = cb.y;
      double t;
      t = dmin(dmax(y[n], minp), maxp);         a0 = fma(y[n] - t, ca.x, a0);     y[n] = t;
      t = dmin(dmax(y[n + 1], minp), maxp);     a1 = fma(y[n + 1] - t, ca.y, a1); y[n + 1] = t;
      t = dmin(dmax(y[n + 2], minp), maxp);     a2 = fma(y[n + 2] - t, cb.x, a2); y[n + 2] = t;
      t = dmin(dmax(y[n + 3], minp), maxp);     a3 = fma(y[n + 3] - t, cb.y, a3); y[n + 3] = t;
    }
    am = (a0 + a1) + (a2 + a3);
  }
  if (fabs(am) > thresh) {
    // From here on the direction is fixed (see above); z = x for addmass > 0, z = -x (bound -minp) for addmass < 0.
    // A node takes part in the redistribution while z < bz.  ce[] carries its weight c while it does and 0 afterwards, so a
    // sweep needs no test "is this node still below the bound": per node
    //     t = z + inc                    (:1060, :1072: x = x + addmass/weightssum; nodes at the bound are put back by the clip)
    //     z = t < bz ? t : bz            (:1037-1040 of the next sweep)
    //     addmass += ce*(t - z)          ((x - maxp)*c for the clipped nodes, 0 for the others and for nodes that were at the bound)
    //     ce = t < bz ? ce : 0;  weightssum += ce
    // which is arithmetically the reference's sequence for every node that moves (t, t - bz and the sums are the same operations
    // on the same operands); 5 FP64-pipe instructions + 4 selects per node.
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
      for (int n = 0; n < 16; n += 4) {
        ce[n] = y[n] < bz ? ce[n] : 0.0;             w0 += ce[n];
        ce[n + 1] = y[n + 1] < bz ? ce[n + 1] : 0.0; w1 += ce[n + 1];
        ce[n + 2] = y[n + 2] < bz ? ce[n + 2] : 0.0; w2 += ce[n + 2];
        ce[n + 3] = y[n + 3] < bz ? ce[n + 3] : 0.0; w3 += ce[n + 3];
      }
      wsum = (w0 + w1) + (w2 + w3);
    }
#pragma unroll 1
    for (int iter = 1; iter <= NPSQ - 1; ++iter) {
      // no node left below the bound: the reference divides by zero, moves nothing and leaves at the next test (:1047)
      if (!(wsum > 0.0)) break;
      const double inc = am / wsum;
      if (iter == NPSQ - 1) {  // the reference's last sweep redistributes without a further clip
        TSE_UNROLL
        for (int n = 0; n < 16; ++n)
          if (y[n] < bz) y[n] += inc;
        break;
      }
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
#define TSE_SWEEP_NODE(n, acc, wacc)     \
  {                                      \
    const double t = y[n] + inc;         \
    const bool below = t < bz;           \
    const double z = below ? t : bz;     \
    acc = fma(ce[n], t - z, acc);        \
    ce[n] = below ? ce[n] : 0.0;         \
    wacc += ce[n];                       \
    y[n] = z;                            \
  }
      TSE_UNROLL
      for (int n = 0; n < 16; n += 4) {
        TSE_SWEEP_NODE(n + 0, a0, w0)
        TSE_SWEEP_NODE(n + 1, a1, w1)
        TSE_SWEEP_NODE(n + 2, a2, w2)
        TSE_SWEEP_NODE(n + 3, a3, w3)
      }
#undef TSE_SWEEP_NODE
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
    const double2 c = lds128v(cbase + cc * PCS);
    y[2 * cc] *= c.x;
    y[2 * cc + 1] *= c.y;
  }
}

// the whole limiter on one plane (check, then the sweeps if needed)
__device__ __forceinline__ void limiter_y(double (&y)[16], unsigned cbase, unsigned rcbase, double sumc, double& minp, double& maxp) {
  if (sumc <= 0.0) return;  // (:1016)
  double mass;
#ifdef TSE_EXP_LIMSTATS  // counting experiment: warps that reach the limiter / that enter the sweeps / lanes active in them
  {
    const bool need = limiter_check(y, rcbase, sumc, minp, maxp, mass);
    const unsigned act = __activemask(), bal = __ballot_sync(act, need);
    if ((threadIdx.x & 31) == (__ffs(act) - 1)) {
      atomicAdd(&g_lim_stats[0], 1ull);
      atomicAdd(&g_lim_stats[1], (unsigned long long)__popc(act));
      if (bal) {
        atomicAdd(&g_lim_stats[2], 1ull);
        atomicAdd(&g_lim_stats[3], (unsigned long long)__popc(bal));
      }
    }
    if (need) limiter_slow(y, cbase, rcbase, mass, minp, maxp);
  }
#elif defined(TSE_EXP_SKIP_SLOW)  // timing experiment only: wrong results (the check runs, the sweeps do not)
  if (limiter_check(y, rcbase, sumc, minp, maxp, mass)) y[0] += 1e-300 * mass;
#else
  if (limiter_check(y, rcbase, sumc, minp, maxp, mass)) limiter_slow(y, cbase, rcbase, mass, minp, maxp);
#endif
}

// Verification hook (tse_debug_limiter): the limiter exactly as the stage kernels call it -- c and 1/c staged in shared memory in
// the package layout, sumc from 4 interleaved partial sums, y = sphweights*ptens in and out -- on independent planes.
__global__ void __launch_bounds__(GPL) k_debug_limiter(int n, double* __restrict__ y, const double* __restrict__ sphweights,
                                                      const double* __restrict__ dpmass, double* __restrict__ minp, double* __restrict__ maxp) {
  __shared__ __align__(16) unsigned char pk[2 * PP_BYTES];  // CL | RC
  const int pl = threadIdx.x, p = blockIdx.x * GPL + pl;
  const bool valid = p < n;
  double yy[16];
  TSE_UNROLL
  for (int c = 0; c < 8; ++c) {
    double2 cl = make_double2(1.0, 1.0), sw = cl;
    if (valid) {
      sw = *reinterpret_cast<const double2*>(sphweights + (size_t)p * 16 + 2 * c);
      const double2 dm = *reinterpret_cast<const double2*>(dpmass + (size_t)p * 16 + 2 * c);
      const double2 pt = *reinterpret_cast<const double2*>(y + (size_t)p * 16 + 2 * c);
      cl = make_double2(sw.x * dm.x, sw.y * dm.y);
      yy[2 * c] = sw.x * pt.x;
      yy[2 * c + 1] = sw.y * pt.y;
    } else {
      yy[2 * c] = yy[2 * c + 1] = 0.0;
    }
    *reinterpret_cast<double2*>(pk + c * PCS + pl * 16) = cl;
    *reinterpret_cast<double2*>(pk + PP_BYTES + c * PCS + pl * 16) = make_double2(1.0 / cl.x, 1.0 / cl.y);
  }
  __syncthreads();
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  TSE_UNROLL
  for (int c = 0; c < 8; c += 2) {
    const double2 x0 = lds128(pk, c * PCS + pl * 16), x1 = lds128(pk, (c + 1) * PCS + pl * 16);
    s0 += x0.x; s1 += x0.y; s2 += x1.x; s3 += x1.y;
  }
  const double sumc = (s0 + s1) + (s2 + s3);
  double mn = valid ? minp[p] : 0.0, mx = valid ? maxp[p] : 0.0;
  const unsigned base = (unsigned)__cvta_generic_to_shared(pk) + pl * 16;
  limiter_y(yy, base, base + PP_BYTES, sumc, mn, mx);
  if (valid) {
    TSE_UNROLL
    for (int c = 0; c < 8; ++c) *reinterpret_cast<double2*>(y + (size_t)p * 16 + 2 * c) = make_double2(yy[2 * c], yy[2 * c + 1]);
    minp[p] = mn;
    maxp[p] = mx;
  }
}

// y(a,b) = sum_i Dvv(i,a) g1(i,b) + sum_i Dvv(i,b) g2(a,i) with g_c = U_c*S, evaluated row by row / column pair by column pair
// so that only S, y and 8 temporaries are live (div_contract needs S, g1, g2 and y at once)
__device__ __forceinline__ void flux_div(const double (&S)[16], const unsigned char* u1, const unsigned char* u2, int pl, const Dvv& D,
                                         double (&y)[16]) {
  TSE_UNROLL
  for (int b = 0; b < 4; ++b) {
    const double2 ua = lds128(u1, (2 * b) * PCS + pl * 16), ub = lds128(u1, (2 * b + 1) * PCS + pl * 16);
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
      const double2 u = lds128(u2, (h + 2 * i) * PCS + pl * 16);
      ge[i] = u.x * S[2 * h + 4 * i];
      go[i] = u.y * S[2 * h + 1 + 4 * i];
    }
    TSE_UNROLL
    for (int b = 0; b < 4; ++b) {
      // the second contraction continues the FMA chain of the first (no separate partial sum and add)
      double se = y[2 * h + 4 * b], so = y[2 * h + 1 + 4 * b];
      TSE_UNROLL
      for (int i = 0; i < 4; ++i) {
        if (!(i == b && (b == 1 || b == 2))) {
          se = fma(D.d[i + 4 * b], ge[i], se);
          so = fma(D.d[i + 4 * b], go[i], so);
        }
      }
      y[2 * h + 4 * b] = se;
      y[2 * h + 1 + 4 * b] = so;
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

// DSS of one level field (the DSSopt variable of euler_step, prim_advection_mod.F90:913-919,943-958):
// out = rspheremp * sum_{sharing elements} spheremp*f, in the reference's unpack order.  CTA = group, walking the 18 level chunks.
// The group's 64 planes of a chunk (8 KB contiguous) are staged in shared memory as spheremp*f with coalesced loads, the nodes
// of the group's perimeter that belong to other groups (the halo list of the tile kernels) or other GPUs (ghost_lev, already
// weighted) next to them; a thread then sums the up to 3 contributions of each of its 4 nodes from shared memory.  Products are
// rounded before they are added (__dmul_rn: no FMA contraction): a neighbour on another GPU arrives as the already rounded
// product, and the sum must be bitwise the same either way.  (The node-per-thread version chased gather-table -> neighbour ->
// value through global memory for 12 of 16 nodes: 1.0 ms per field at ne120, a quarter of the HBM rate.)
__constant__ signed char c_node_slots_t[16][3] = {
    {0, 12, 16}, {1, -1, -1}, {2, -1, -1},  {3, 4, 17},   // j = 0: S edge; node 0 also W and SW, node 3 also E and SE
    {13, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {5, -1, -1},  // j = 1: W, interior, interior, E
    {14, -1, -1}, {-1, -1, -1}, {-1, -1, -1}, {6, -1, -1},  // j = 2
    {8, 15, 19}, {9, -1, -1}, {10, -1, -1}, {7, 11, 18}};  // j = 3: N edge; node 12 = N, W, NW; node 15 = E, N, NE
constexpr int DSL_THREADS = GPL * 4;  // 4 threads per plane, 4 consecutive nodes (one row) each
__host__ __device__ constexpr int dss_level_smem_bytes(int hmax) { return 2 * (GPL * 16 + hmax * KC) * 8; }

__global__ void __launch_bounds__(DSL_THREADS) k_dss_level(Geo G, TileTables tb, const double* __restrict__ f, const double* __restrict__ ghost,
                                                          double* __restrict__ out) {
  extern __shared__ double dsm[];  // [buf 2][GPL*16 tile | hmax*KC halo]
  const int t = threadIdx.x, g = blockIdx.x;
  const int W = GPL * 16 + tb.hmax * KC;
  const int hoff = tb.halo_off[g], H = tb.halo_off[g + 1] - hoff;
  // loader role: 4 consecutive nodes of plane pl (coalesced 32-byte pieces); the halo entries (h, kk), level fastest
  const int pl = t >> 2, r4 = (t & 3) * 4, el = pl / KC, kk = pl % KC;
  const int e = min(g * GE + el, G.nelem - 1);
  const double4 sp = *reinterpret_cast<const double4*>(G.spheremp + (size_t)e * 16 + r4);
  const double4 rs = *reinterpret_cast<const double4*>(G.rspheremp + (size_t)e * 16 + r4);
  constexpr int HI = 2;  // halo items per thread in registers (hmax*KC <= 2*DSL_THREADS = 512 covers hmax <= 128)
  const double* h_src[HI];
  double h_w[HI];
  int h_stride[HI];
  TSE_UNROLL
  for (int i = 0; i < HI; ++i) {
    const int it = t + i * DSL_THREADS;
    h_src[i] = f; h_w[i] = 0.0; h_stride[i] = 0;
    if (it < H * KC) {
      const int h = it / KC, k2 = it % KC;
      const int code = tb.halo_src[hoff + h];
      if (code >= 0) {  // node (code & 15) of element (code >> 4): spheremp * f
        const int es = code >> 4, nd = code & 15;
        h_src[i] = f + lplane(es, k2) * 16 + nd;
        h_w[i] = G.spheremp[(size_t)es * 16 + nd];
        h_stride[i] = GPL * 16;  // doubles between level chunks of a level field
      } else {          // ghost slot: already the rounded product, [slot][k]
        h_src[i] = ghost + (size_t)(-code - 2) * NLEV + k2;
        h_w[i] = -1.0;
        h_stride[i] = KC;
      }
    }
  }
  // consumer role: the same 4 nodes; up to 3 contributions each, as offsets into a shared-memory buffer (-1: none)
  short coff[4][3];
  {
    const int* gs = tb.gsrc_t + (size_t)e * NSLOT;
    TSE_UNROLL
    for (int j = 0; j < 4; ++j)
      TSE_UNROLL
      for (int c = 0; c < 3; ++c) {
        const int slot = c_node_slots_t[r4 + j][c];
        int off = -1;
        if (slot >= 0) {
          const int code = gs[slot];
          if (code >= 256) off = GPL * 16 + (code - 256) * KC + kk;
          else if (code >= 0) off = ((code >> 4) * KC + kk) * 16 + (code & 15);
        }
        coff[j][c] = (short)off;
      }
  }
  const bool evalid = g * GE + el < G.nelem;
  const size_t base = ((size_t)g * NKC * GPL + pl) * 16 + r4;  // (chunk 0, plane pl, node r4) of this group in a level field
  double4 r_own;
  double r_h[HI];
  auto prefetch = [&](int kc) {
    r_own = *reinterpret_cast<const double4*>(f + base + (size_t)kc * GPL * 16);
    TSE_UNROLL
    for (int i = 0; i < HI; ++i)
      if (t + i * DSL_THREADS < H * KC) r_h[i] = h_src[i][(size_t)kc * h_stride[i]];
  };
  prefetch(0);
  for (int kc = 0; kc < NKC; ++kc) {
    double* sb = dsm + (size_t)(kc & 1) * W;
    *reinterpret_cast<double4*>(sb + pl * 16 + r4) =
        make_double4(__dmul_rn(sp.x, r_own.x), __dmul_rn(sp.y, r_own.y), __dmul_rn(sp.z, r_own.z), __dmul_rn(sp.w, r_own.w));
    TSE_UNROLL
    for (int i = 0; i < HI; ++i)
      if (t + i * DSL_THREADS < H * KC) sb[GPL * 16 + t + i * DSL_THREADS] = h_w[i] < 0.0 ? r_h[i] : __dmul_rn(h_w[i], r_h[i]);
    for (int it = t + HI * DSL_THREADS; it < H * KC; it += DSL_THREADS) {  // groups with more than 128 perimeter nodes
      const int h = it / KC, k2 = it % KC;
      const int code = tb.halo_src[hoff + h];
      const int kq = kc * KC + k2;
      sb[GPL * 16 + it] = code >= 0 ? __dmul_rn(G.spheremp[(size_t)(code >> 4) * 16 + (code & 15)], f[lplane(code >> 4, kq) * 16 + (code & 15)])
                                    : ghost[(size_t)(-code - 2) * NLEV + kq];
    }
    __syncthreads();
    if (kc + 1 < NKC) prefetch(kc + 1);
    const double* own = sb + pl * 16 + r4;
    double v[4] = {own[0], own[1], own[2], own[3]};
    TSE_UNROLL
    for (int j = 0; j < 4; ++j)
      TSE_UNROLL
      for (int c = 0; c < 3; ++c)
        if (coff[j][c] >= 0) v[j] += sb[coff[j][c]];
    if (evalid)
      *reinterpret_cast<double4*>(out + base + (size_t)kc * GPL * 16) = make_double4(v[0] * rs.x, v[1] * rs.y, v[2] * rs.z, v[3] * rs.w);
  }
}

// min/max over the element and its up to 8 neighbours (neighbor_minmax, viscosity_mod.F90:748-816).
// CTA = (group, level chunk), walking the tracers two at a time.  The element extrema of the group (2 x 512 contiguous bytes per
// tracer) and of the neighbour elements outside the group (one 32-byte sector each: the KC = 4 levels of a chunk are contiguous
// both in the per-plane arrays and in the ghost bundles) are staged in shared memory with 16-byte loads; each thread then forms
// the 9-way min/max of one (element, level) from shared memory.  The plane-per-thread version read every neighbour straight
// from global memory, 18 sector requests per plane, and ran into the L1 wavefront limit (92 % of the LSU data pipe, 0.36 of the
// HBM roofline); here a sector is requested once per group.
struct NbrTables {
  const int* nbr_t;    // [npad][8]: < 0 none, [0, GE) element of the same group, >= 256: external entry (code - 256) of the group
  const int* ext_off;  // [ngroups + 1]
  const int* ext_src;  // [ext_off[ngroups]]: >= 0 element (internal order), <= -2 ghost bundle -(v + 2)
  int xmax;            // max external entries of a group
};
#ifndef TSE_NBQ
#define TSE_NBQ 2
#endif
constexpr int NBQ = TSE_NBQ;  // tracers per batch = NBQ * GPL threads
__host__ __device__ constexpr int nbr_smem_bytes(int xmax) { return 2 * 2 * NBQ * (GPL + xmax * KC) * 8; }
static_assert(KC == 4 && NLEV % KC == 0, "k_nbr_minmax moves the 4 levels of a chunk as two double2");

__global__ void __launch_bounds__(NBQ* GPL) k_nbr_minmax(Geo G, NbrTables nt, int Q, const double* __restrict__ lmin,
                                                        const double* __restrict__ lmax, const double* __restrict__ ghost_mm,
                                                        double* __restrict__ qmin, double* __restrict__ qmax) {
  extern __shared__ double nsm[];  // [buf 2][arr 2][q NBQ][GPL + xmax*KC]
  const int t = threadIdx.x;
  const int ngl = gridDim.x / NKC;
  const int g = blockIdx.x % ngl, kc = blockIdx.x / ngl;  // chunk-major like k_pipe: neighbouring groups run together (L2)
  const int W = GPL + nt.xmax * KC;                        // scalars per (array, tracer) in shared memory
  const int xo = nt.ext_off[g], nx = nt.ext_off[g + 1] - xo;
  const size_t tile0 = (((size_t)g * NKC + kc) * Q) * GPL;  // plane index of (tracer 0, plane 0) of this CTA

  // loader role.  Own tiles: NBQ*2 arrays x GPL scalars = one double2 per thread.  External sectors: item = (q, arr, x, half).
  const int own_arr = (t / (GPL / 2)) & 1, own_q = t / GPL, own_h = t % (GPL / 2);
  constexpr int XI = 2;  // external items per thread (covers xmax <= 32 with NBQ*GPL = 128 threads; more loop below)
  // consumer role: one (tracer, element, level) per thread
  const int cq = t / GPL, pl = t % GPL, el = pl / KC, kk = pl % KC;
  const int e = g * GE + el;
  int noff[8];
  {
    const int4* n4 = reinterpret_cast<const int4*>(nt.nbr_t + (size_t)min(e, G.nelem - 1) * 8);
    const int4 u = n4[0], v = n4[1];
    const int c[8] = {u.x, u.y, u.z, u.w, v.x, v.y, v.z, v.w};
    TSE_UNROLL
    for (int d = 0; d < 8; ++d) noff[d] = c[d] < 0 ? pl : c[d] >= 256 ? GPL + (c[d] - 256) * KC + kk : c[d] * KC + kk;
  }
  const int nitems = 2 * 2 * NBQ * nx;  // (half, x, arr, q)
  // decode an external item once: source of tracer slot 0 .. NBQ-1 of batch 0 (pointer to tracer ql, stride per tracer) and
  // the destination in a shared-memory buffer
  auto ext_decode = [&](int it, const double*& base, int& stride, int& ql, int& dst) {
    const int half = it & 1, idx = it >> 1, x = idx % nx, r = idx / nx, arr = r & 1;
    ql = r >> 1;
    const int src = nt.ext_src[xo + x];
    if (src >= 0) {
      base = (arr ? lmax : lmin) + qplane(src, 0, kc * KC, Q) + 2 * half;
      stride = GPL;
    } else {
      base = ghost_mm + ((size_t)(-src - 2) * 2 + arr) * Q * NLEV + kc * KC + 2 * half;
      stride = NLEV;
    }
    dst = (arr * NBQ + ql) * W + GPL + x * KC + 2 * half;
  };
  const double* e_base[XI];
  int e_stride[XI], e_ql[XI], e_dst[XI];
  TSE_UNROLL
  for (int i = 0; i < XI; ++i) {
    e_base[i] = lmin; e_stride[i] = 0; e_ql[i] = 0; e_dst[i] = 0;
    const int it = t + i * NBQ * GPL;
    if (it < nitems) ext_decode(it, e_base[i], e_stride[i], e_ql[i], e_dst[i]);
  }
  const double* const own_base = (own_arr ? lmax : lmin) + tile0 + 2 * own_h;
  const int nb = (Q + NBQ - 1) / NBQ;
  double2 r_own, r_ext[XI];
  auto prefetch = [&](int b) {
    const int q0 = b * NBQ;
    r_own = *reinterpret_cast<const double2*>(own_base + (size_t)min(q0 + own_q, Q - 1) * GPL);
    TSE_UNROLL
    for (int i = 0; i < XI; ++i)
      if (t + i * NBQ * GPL < nitems) r_ext[i] = *reinterpret_cast<const double2*>(e_base[i] + (size_t)min(q0 + e_ql[i], Q - 1) * e_stride[i]);
  };
  prefetch(0);
  for (int b = 0; b < nb; ++b) {
    double* sb = nsm + (size_t)(b & 1) * 2 * NBQ * W;
    *reinterpret_cast<double2*>(sb + (own_arr * NBQ + own_q) * W + 2 * own_h) = r_own;
    TSE_UNROLL
    for (int i = 0; i < XI; ++i)
      if (t + i * NBQ * GPL < nitems) *reinterpret_cast<double2*>(sb + e_dst[i]) = r_ext[i];
    for (int it = t + XI * NBQ * GPL; it < nitems; it += NBQ * GPL) {  // groups with more than 32 external neighbours
      const double* base;
      int stride, ql, dst;
      ext_decode(it, base, stride, ql, dst);
      *reinterpret_cast<double2*>(sb + dst) = *reinterpret_cast<const double2*>(base + (size_t)min(b * NBQ + ql, Q - 1) * stride);
    }
    __syncthreads();
    if (b + 1 < nb) prefetch(b + 1);
    const int q = b * NBQ + cq;
    const double* smn = sb + cq * W;
    const double* smx = sb + (NBQ + cq) * W;
    double mn = smn[pl], mx = smx[pl];
    TSE_UNROLL
    for (int d = 0; d < 8; ++d) {
      mn = dmin(mn, smn[noff[d]]);
      mx = dmax(mx, smx[noff[d]]);
    }
    if (q < Q && e < G.nelem) {
      const size_t pidx = tile0 + (size_t)q * GPL + pl;
      qmin[pidx] = mn;
      qmax[pidx] = mx;
    }
  }
}

}  // namespace tse
