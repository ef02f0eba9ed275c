// Tiled tracer-field kernels (sm_100a), warp-specialised: the element operators, limiter, package and tile layout are in
// tse_tile.cuh; this file is the pipeline around them.
//
//   producer warps  (2) move everything that comes from HBM.  Per pipeline item (QI tracers of one input field) one lane issues
//                   the tile as TMA tensor copies (cp.async.bulk.tensor.2d, 16 planes x 128 B per box, SWIZZLE_128B = the XOR
//                   swizzle the plane-per-thread reads need) and, for the stage ops, the limiter bounds of the item's planes as
//                   two 1-D bulk copies; all lanes issue the DSS halo (68 x KC scattered nodes per tracer for a 4x4 patch) as
//                   8-byte cp.async.  Everything completes on the stage's "full" mbarrier (complete_tx bytes /
//                   cp.async.mbarrier.arrive.noinc).  The producers run up to NST items ahead of the math and never touch
//                   registers the math needs.
//   consumer warps  (TT threads, one 4x4 plane per thread) wait on "full", pull their plane and its DSS neighbours into
//                   registers, release the stage at once ("empty" mbarrier, one arrive per thread), do the arithmetic, stage
//                   the result in the OUT tile and hand it to a TMA store (one 2 KB box per tracer per warp, bulk async
//                   group per warp).  There is no CTA-wide barrier in the steady state: a warp that sits in a long limiter
//                   loop delays nobody until the pipeline runs dry.
//
// Rows of the tensor maps are planes (16 doubles) of a tracer field in layout order, so a tile is a run of consecutive rows.
#pragma once
#include <cuda.h>

#include "tse_tile.cuh"

namespace tse {

// IN stages / OUT buffers per op.  The streaming ops (little math per plane) are bound by bytes in flight: with 2 stages a
// CTA moves 32 KB per release -> refill round trip (about 2 us), i.e. 4.7 TB/s over the chip; they get 3-4 stages and, where
// it still leaves 2 CTAs per SM (113 KB each), a double-buffered OUT tile.  The stage ops carry a 40 KB package: 2 + 1.
// (OP_MINMAX measured slower with 4 stages than with 2.)
#ifndef TSE_BIHARM3
#define TSE_BIHARM3 0  // experiment: OP_BIHARM_PRE with 2 stages and 3 CTAs per SM (112 registers)
#endif
__host__ __device__ constexpr int pipe_nst(int op) {
  return (op == OP_BIHARM_PRE && TSE_BIHARM3) ? 2 : (op == OP_BIHARM_PRE || op == OP_TIME_AVG || op == OP_RESOLVE || op == OP_MASS) ? 3 : 2;
}
__host__ __device__ constexpr int pipe_minb(int op) { return (op == OP_BIHARM_PRE && TSE_BIHARM3) ? 3 : TSE_MINB; }
__host__ __device__ constexpr int pipe_nout(int op) { return (op == OP_TIME_AVG || op == OP_RESOLVE) ? 2 : 1; }
constexpr int NST_MAX = 4;               // barrier slots reserved per kind
// Stage release protocol (consumer -> producer, "this stage may be refilled"):
//   0  fence.proxy.async.shared::cta, then mbarrier.arrive.  The refill is written by the async proxy (TMA, cp.async.bulk);
//      the consumer's reads are generic-proxy ld.shared.  mbarrier release/acquire orders generic-proxy accesses only, so the
//      write-after-read across the two proxies needs the proxy fence on the reading side (PTX ISA "async proxy"; the same
//      fence CUTLASS issues before consumer_release when a TMA load overwrites a buffer read with ld.shared).
//   1  round-1 protocol: the arrive count is made data-dependent on every loaded register, so the arrive cannot issue before
//      the loads have returned.  Correct on the hardware, but outside the memory model.
//   2  plain arrive without fence: WRONG (rare stale planes at ne >= 90 on the B200); kept for the sanitizer runs.
#ifndef TSE_RELEASE
#define TSE_RELEASE 0
#endif
constexpr int NCW = TT / 32;             // consumer warps
constexpr int MASS_REP = 64;             // copies of the OP_MASS accumulators
#ifndef TSE_NPW
#define TSE_NPW 2
#endif
constexpr int NPW = TSE_NPW;             // producer warps: 2 (the 8-byte halo gathers are issue-bound in one warp: 129 -> 124 ms per
                                         // tracer step at ne120; 4 warps with a setmaxnreg 40/216 register split measured no better)
// producer warps / threads per CTA of an op (OP_MINMAX has no halo in the time loop: one producer, and 3 CTAs fit per SM)
__host__ __device__ constexpr int pipe_npw(int op) { return op == OP_MINMAX ? 1 : NPW; }
__host__ __device__ constexpr int pipe_threads(int op) { return TT + 32 * pipe_npw(op); }
constexpr int BOX_ROWS = EPW * KC;       // planes per TMA box = one warp's planes of one tracer
static_assert((BOX_ROWS == 8 || BOX_ROWS == 16) && GPL % BOX_ROWS == 0, "TMA box = whole 1 KB swizzle atoms (8 planes)");
static_assert(EPW > 1, "SWIZZLE_128B is the row&7 XOR");

__host__ __device__ constexpr int pipe_in_stride(int hmax) { return (tile_in_bytes(hmax) + 1023) & ~1023; }  // (stage ops use the bounds area)
__host__ __device__ constexpr int pipe_smem_bytes(int op, int hmax) {
  return 1024 + pipe_nst(op) * pipe_in_stride(hmax) + (tile_cfg(op).has_out ? pipe_nout(op) * TILE_BYTES : 0) +
         tile_cfg(op).npp * PP_BYTES + tile_cfg(op).nel * EL_BYTES + hmax * KC * 10 + 16 + 2 * NST_MAX * 8;
}

struct PipeMaps {
  CUtensorMap in[2];  // the two input fields (box = BOX_ROWS planes)
  CUtensorMap out;
};

__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(addr) : "memory");
}
// arrive with an explicit count (always 1): `dep` is OR-ed in through a runtime zero, which makes the arrive wait for the
// registers `dep` was computed from (see the stage release in k_pipe)
__device__ __forceinline__ void mbar_arrive_after(unsigned addr, unsigned dep, unsigned zero) {
  const unsigned cnt = 1u | (dep & zero);
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;\n" ::"r"(addr), "r"(cnt) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned addr, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
#ifndef TSE_BACKOFF_NS
#define TSE_BACKOFF_NS 256
#endif
// same, for waits that are expected to be long (the producer waiting for a stage to drain): back off between polls so that
// the spinning warp does not take issue slots from the math warps of its scheduler
__device__ __forceinline__ void mbar_wait_backoff(unsigned addr, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "LAB_WAIT:\n"
      "nanosleep.u32 %2;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\n"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n"
      "}\n" ::"r"(addr),
      "r"(parity), "n"(TSE_BACKOFF_NS)
      : "memory");
}
// completion of all cp.async issued so far by this thread counts as one (pre-counted) arrival on the mbarrier
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(unsigned addr) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned mbar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(mbar)
               : "memory");
}
// 1-D bulk copy global -> shared (bytes a multiple of 16), completing on an mbarrier
__device__ __forceinline__ void bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, unsigned src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];\n" ::"l"(map), "r"(c0), "r"(c1), "r"(src)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;\n" ::"n"(TT) : "memory"); }

template <int OP>
__global__ void __launch_bounds__(pipe_threads(OP), (pipe_threads(OP) > 256 ? 1 : pipe_minb(OP))) k_pipe(const __grid_constant__ PipeMaps maps, Geo G, Dvv D, TileTables tb, TileArgs a) {
  constexpr TileCfg cfg = tile_cfg(OP);
  constexpr bool kS3 = (OP == OP_STAGE3 || OP == OP_HYPERVIS);
  constexpr bool kStage = (OP == OP_STAGE1 || OP == OP_STAGE2 || kS3);
  constexpr int NIN = (kS3 || OP == OP_TIME_AVG) ? 2 : 1;
  constexpr bool kHasOut = cfg.has_out != 0;
  constexpr int NST = pipe_nst(OP), NOUT = pipe_nout(OP), NPW = pipe_npw(OP);
  extern __shared__ unsigned char smem_raw[];
  const unsigned raw_u32 = (unsigned)__cvta_generic_to_shared(smem_raw);
  unsigned char* const smem = smem_raw + ((1024u - (raw_u32 & 1023u)) & 1023u);  // SWIZZLE_128B tiles sit on 1 KB boundaries
  const unsigned smem_u32 = (unsigned)__cvta_generic_to_shared(smem);
  const int IN_STRIDE = pipe_in_stride(tb.hmax);
  unsigned char* const outb = smem + NST * IN_STRIDE;
  unsigned char* const pp = outb + (kHasOut ? NOUT * TILE_BYTES : 0);
  unsigned char* const elb = pp + cfg.npp * PP_BYTES;
  long long* const htab = reinterpret_cast<long long*>(elb + cfg.nel * EL_BYTES);  // halo sources (double index, tracer 0)
  unsigned short* const hdtab = reinterpret_cast<unsigned short*>(htab + tb.hmax * KC);  // halo destinations (8-byte units)
  const unsigned bar_u32 = (smem_u32 + (unsigned)(reinterpret_cast<unsigned char*>(hdtab + tb.hmax * KC) - smem) + 15u) & ~15u;
  const int ZERO_OFF = TILE_BYTES + QI * tb.hmax * KC * 8;
  const int BND_OFF = ZERO_OFF + 16;  // limiter bounds of the item: qmin[QI*GPL], qmax[QI*GPL] (stage ops)
  auto full_bar = [&](int b) -> unsigned { return bar_u32 + b * 8; };
  auto empty_bar = [&](int b) -> unsigned { return bar_u32 + (NST + b) * 8; };

  const int t = threadIdx.x;
  // CTAs are numbered level chunk by level chunk: the ones in flight together then cover one compact patch of the sphere, and
  // most halo nodes (which belong to groups far away along the space-filling curve) are found in L2 instead of HBM
  // (measured at ne120: 7-13 % less DRAM traffic than group-major numbering)
  const int ngl = gridDim.x / NKC;
  const int gi = blockIdx.x % ngl, kc = blockIdx.x / ngl;
  const int g = a.glist ? a.glist[gi] : gi;
  const int w = t >> 5, lane = t & 31;
  const int Q = a.Q;
  const int nit = (Q + QI - 1) / QI;
  const int nitems = nit * NIN;
  const unsigned row0 = (unsigned)(((size_t)g * NKC + kc) * Q * GPL);  // tensor-map row of (tracer 0, plane 0) of this CTA

  if (t == 0) {
    for (int b = 0; b < NST; ++b) {
      mbar_init(full_bar(b), 1 + 32 * NPW);  // expect_tx arrive of the issuing lane + one cp.async arrival per producer lane
      mbar_init(empty_bar(b), TT);  // every consumer thread releases for itself
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (t < NST) *reinterpret_cast<double*>(smem + t * IN_STRIDE + ZERO_OFF) = 0.0;
  fence_proxy_async_smem();
  __syncthreads();

  if (w >= NCW) {
    const int pw = w - NCW;  // producer warp index: warp 0 issues the tiles, all share the halo

    // =============================================== producer warp ===============================================
    const bool any_pending = a.pending[0] || (NIN == 2 && a.pending[1]);
    auto issue_tile = [&](int j) {  // TMA tile of item j into its stage (after the consumers have drained it)
      const int b = j % NST, it = j / NIN, which = j % NIN;
      const int q0 = it * QI, nq = min(QI, Q - q0);
      if (j >= NST) mbar_wait_backoff(empty_bar(b), (unsigned)(((j / NST) - 1) & 1));
      if (lane == 0 && pw == 0) {
        const unsigned sb = smem_u32 + b * IN_STRIDE;
        const bool bounds = kStage && which == NIN - 1;  // the item the limiter runs on also brings its bounds (2 x nq*GPL doubles)
        mbar_arrive_expect_tx(full_bar(b), (unsigned)(nq * GPL * 128 + (bounds ? 2 * nq * GPL * 8 : 0)));
        const CUtensorMap* m = &maps.in[which];
        const int r0 = (int)(row0 + (unsigned)q0 * GPL);
        for (int bx = 0; bx < nq * (GPL / BOX_ROWS); ++bx) tma_load_2d(sb + bx * (BOX_ROWS * 128), m, 0, r0 + bx * BOX_ROWS, full_bar(b));
        if (bounds) {
          const size_t off = (size_t)row0 + (size_t)q0 * GPL;  // plane index of (tracer q0, plane 0): bounds are per plane
          bulk_load(sb + BND_OFF, a.qmin + off, (unsigned)(nq * GPL * 8), full_bar(b));
          bulk_load(sb + BND_OFF + QI * GPL * 8, a.qmax + off, (unsigned)(nq * GPL * 8), full_bar(b));
        }
      }
    };
    int nhalo = 0;
    auto issue_halo = [&](int j) {  // DSS halo of item j; every lane then arrives on the stage's full barrier
      const int b = j % NST, it = j / NIN, which = j % NIN;
      const int q0 = it * QI, nq = min(QI, Q - q0);
      if (a.pending[which]) {
        const unsigned sb = smem_u32 + b * IN_STRIDE;
        const double* src = a.src[which];
        const double* ghost = a.ghost[which];
        for (int idx = lane + 32 * pw; idx < nhalo; idx += 32 * NPW) {
          const long long v = htab[idx];
          const unsigned dst = sb + ((unsigned)hdtab[idx] << 3);
          for (int qi2 = 0; qi2 < nq; ++qi2) {
            const double* gp = v >= 0 ? src + v + (size_t)(q0 + qi2) * GPL * 16 : ghost + (-(v + 1)) + (size_t)(q0 + qi2) * NLEV;
            cp_async8(dst + qi2 * tb.hmax * KC * 8, gp);
          }
        }
      }
      cp_async_mbar_arrive_noinc(full_bar(b));
    };
    // the first tiles go out before anything else: they need no table
    const int nfirst = min(NST, nitems);
    for (int j = 0; j < nfirst; ++j) issue_tile(j);
    if (any_pending) {
      const int hoff = tb.halo_off[g], H = tb.halo_off[g + 1] - hoff;
      nhalo = H * KC;
      // entry idx -> (h = idx / KC, kk2 = idx % KC), level fastest: consecutive lanes then write consecutive 8-byte slots of the
      // halo array (with the halo node fastest they were 32 bytes apart: 8-way bank conflicts, 11 wavefronts per LDGSTS, 10-18 % of
      // the shared-memory wavefronts of the ops that gather) and read the 4 levels of one node, 4 sectors of consecutive lines
      for (int idx = lane + 32 * pw; idx < nhalo; idx += 32 * NPW) {
        const int h = idx / KC, kk2 = idx % KC;
        const int code = tb.halo_src[hoff + h];
        const int kq = kc * KC + kk2;
        htab[idx] = code >= 0 ? (long long)(qplane(code >> 4, 0, kq, Q) * 16 + (code & 15)) : -((long long)(-code - 2) * Q * NLEV + kq) - 1;
        hdtab[idx] = (unsigned short)((TILE_BYTES + (h * KC + kk2) * 8) >> 3);
      }
      if (NPW > 1) asm volatile("bar.sync 2, %0;\n" ::"n"(32 * NPW) : "memory"); else __syncwarp();
    }
    for (int j = 0; j < nfirst; ++j) issue_halo(j);
    for (int j = nfirst; j < nitems; ++j) {
      issue_tile(j);
      issue_halo(j);
    }
    return;
  }

  // ================================================= consumer warps =================================================
  const int kk = lane & 3, el = EPW * (w % (GE / EPW)) + ((lane >> 2) % EPW), qi = QW * (w / (GE / EPW)) + (lane >> 2) / EPW;
  const int pl = el * KC + kk;   // plane within one tracer's tile
  const int p = qi * GPL + pl;   // plane within the QI-tracer tile
  const int e = g * GE + el, k = kc * KC + kk;
  const bool evalid = e < G.nelem;

  // the DSS gather table of this thread's element: loaded first so that its latency overlaps the package loads
  int gsv[NSLOT];
  {
    const int4* gs4 = reinterpret_cast<const int4*>(tb.gsrc_t + (size_t)(evalid ? e : 0) * NSLOT);
    TSE_UNROLL
    for (int s4 = 0; s4 < NSLOT / 4; ++s4) {
      const int4 v = gs4[s4];
      gsv[4 * s4] = v.x; gsv[4 * s4 + 1] = v.y; gsv[4 * s4 + 2] = v.z; gsv[4 * s4 + 3] = v.w;
    }
  }

  // ---- level package (see tse_tile.cuh) ------------------------------------------------------------------------------
  // All global loads of the prologue are issued before the first of them is used (no branches in between: padded elements
  // read the last real element, their planes are never computed): one memory round trip instead of six dependent ones.
  // With 2 CTAs per SM the consumer warps of a starting CTA otherwise sit out ~25 % of the CTA's lifetime here.
  const bool main_pending = kS3 ? (a.pending[1] != 0) : (a.pending[0] != 0);
  const int elast = G.nelem - 1;
  // The stage ops need 11 double2 per (plane, chunk) pair, 6 of them geometry of the element (spheremp, rspheremp, the metric):
  // the group's geometry (12 KB) goes through the OUT tile, which nothing uses before the first item is done, as 16-byte cp.async
  // -- 6 per thread, no registers -- and the 5 level fields of all NPK pairs are loaded at once: one HBM round trip for the
  // whole prologue (with everything in registers it took two batches, i.e. two dependent round trips).
  constexpr bool kGeoStage = kStage && TT == GE * 8 && GPL * 8 == 4 * TT;
  constexpr int GS_SP = 0, GS_RS = GE * 128, GS_MD = 2 * GE * 128;  // [el][16], [el][16], [el][4][16] doubles from outb
  if (kGeoStage) {
    const unsigned ob32 = smem_u32 + (unsigned)(outb - smem);
    const int ee = min(g * GE + (t >> 3), elast), c = t & 7;
    cp_async16(ob32 + GS_SP + t * 16, G.spheremp + (size_t)ee * 16 + 2 * c);
    cp_async16(ob32 + GS_RS + t * 16, G.rspheremp + (size_t)ee * 16 + 2 * c);
    TSE_UNROLL
    for (int j = 0; j < 4; ++j) {
      const int id = t + TT * j, em = min(g * GE + (id >> 5), elast);
      cp_async16(ob32 + GS_MD + id * 16, G.mD + (size_t)em * 64 + 2 * (id & 31));
    }
  }
  double2 L_sp = make_double2(0, 0), L_rs = L_sp, L_rm = L_sp, L_t11 = L_sp, L_t12 = L_sp, L_t22 = L_sp;
  const bool el_thread = cfg.nel > 0 && t < GE * 8;
  if (el_thread) {
    const int c = t & 7, ee = min(g * GE + (t >> 3), elast);
    const size_t b = (size_t)ee * 16 + 2 * c;
    if (!kGeoStage) {
      L_sp = *reinterpret_cast<const double2*>(G.spheremp + b);
      L_rs = *reinterpret_cast<const double2*>(G.rspheremp + b);
    }
    L_rm = *reinterpret_cast<const double2*>(G.rmr + b);
    if (cfg.T11 >= 0) {
      const double* T = G.T + (size_t)ee * 48 + 2 * c;
      L_t11 = *reinterpret_cast<const double2*>(T);
      L_t12 = *reinterpret_cast<const double2*>(T + 16);
      L_t22 = *reinterpret_cast<const double2*>(T + 32);
    }
  }
  // (plane, 16-byte chunk) pairs of the GPL x 8 chunks, strided over the consumer threads; the loads of PB pairs are issued
  // before the first of them is used
  constexpr int NPK = (GPL * 8) / TT, PB = (kGeoStage || !kStage) ? NPK : (NPK % 2 == 0 ? 2 : 1);
  static_assert((GPL * 8) % TT == 0, "package pairs divide evenly over the consumer threads");
  auto store_el = [&]() {
    if (el_thread) {
      const int pe = t >> 3, c = t & 7;
      if (kGeoStage) {
        L_sp = lds128(outb, GS_SP + t * 16);
        L_rs = lds128(outb, GS_RS + t * 16);
      }
      const double rx0 = main_pending ? L_rs.x : 1.0, rx1 = main_pending ? L_rs.y : 1.0;
      const double2 e1 = OP == OP_MASS ? L_sp : make_double2(L_sp.x * rx0, L_sp.y * rx1);
      const double2 e2 = make_double2(a.dt * (L_sp.x * L_rm.x), a.dt * (L_sp.y * L_rm.y));
      const int off = (c * GE + pe) * 16;
      if (cfg.E1 >= 0) *reinterpret_cast<double2*>(elb + cfg.E1 * EL_BYTES + off) = e1;
      if (cfg.E2 >= 0) *reinterpret_cast<double2*>(elb + cfg.E2 * EL_BYTES + off) = e2;
      if (cfg.RSPH >= 0) *reinterpret_cast<double2*>(elb + cfg.RSPH * EL_BYTES + off) = L_rs;
      if (cfg.T11 >= 0) {
        *reinterpret_cast<double2*>(elb + cfg.T11 * EL_BYTES + off) = L_t11;
        *reinterpret_cast<double2*>(elb + cfg.T12 * EL_BYTES + off) = L_t12;
        *reinterpret_cast<double2*>(elb + cfg.T22 * EL_BYTES + off) = L_t22;
      }
    }
  };
  if (cfg.npp == 0) store_el();
  if (cfg.npp > 0) {
    TSE_UNROLL
    for (int r0b = 0; r0b < NPK; r0b += PB) {
      double2 P_dp[PB], P_dj[PB], P_rs[PB], P_dd[PB], P_v1[PB], P_v2[PB], P_sp[PB], P_m11[PB], P_m12[PB], P_m21[PB], P_m22[PB];
      TSE_UNROLL
      for (int r = 0; r < PB; ++r) {
        const int i = t + (r0b + r) * TT;
        const int ppl = i >> 3, n = 2 * (i & 7);
        const int pe = min(g * GE + ppl / KC, elast), pk = kc * KC + ppl % KC;
        const size_t lp = lplane(pe, pk) * 16 + n, gb = (size_t)pe * 16 + n;
        P_dp[r] = *reinterpret_cast<const double2*>(a.dp + lp);
        P_dj[r] = *reinterpret_cast<const double2*>(a.divdp_proj + lp);
        if (!kGeoStage) P_rs[r] = *reinterpret_cast<const double2*>(G.rspheremp + gb);
        if (kStage) {
          P_dd[r] = *reinterpret_cast<const double2*>(a.divdp + lp);
          P_v1[r] = *reinterpret_cast<const double2*>(a.vn0 + vplane(pe, pk, 0) * 16 + n);
          P_v2[r] = *reinterpret_cast<const double2*>(a.vn0 + vplane(pe, pk, 1) * 16 + n);
          if (!kGeoStage) {
            P_sp[r] = *reinterpret_cast<const double2*>(G.spheremp + gb);
            const double* mD = G.mD + (size_t)pe * 64 + n;
            P_m11[r] = *reinterpret_cast<const double2*>(mD);
            P_m12[r] = *reinterpret_cast<const double2*>(mD + 16);
            P_m21[r] = *reinterpret_cast<const double2*>(mD + 32);
            P_m22[r] = *reinterpret_cast<const double2*>(mD + 48);
          }
        }
      }
      if (r0b == 0) {
        if (kGeoStage) {
          cp_async_wait_all();
          consumer_barrier();  // every thread's part of the geometry has landed
        }
        store_el();
      }
      TSE_UNROLL
      for (int r = 0; r < PB; ++r) {
        const int i = t + (r0b + r) * TT;
        const int ppl = i >> 3, c = i & 7;
        if (kGeoStage) {
          const int ge = (ppl / KC) * 128 + c * 16;
          P_sp[r] = lds128(outb, GS_SP + ge);
          P_rs[r] = lds128(outb, GS_RS + ge);
          const int gm = GS_MD + (ppl / KC) * 512 + c * 16;
          P_m11[r] = lds128(outb, gm);
          P_m12[r] = lds128(outb, gm + 128);
          P_m21[r] = lds128(outb, gm + 256);
          P_m22[r] = lds128(outb, gm + 384);
        }
        double2 u1 = make_double2(0, 0), u2 = u1, cl = make_double2(1, 1), rcl = make_double2(1, 1);
        const double rx0 = main_pending ? P_rs[r].x : 1.0, rx1 = main_pending ? P_rs[r].y : 1.0;
        const double dps0 = P_dp[r].x - a.rhs_mult_dt * P_dj[r].x, dps1 = P_dp[r].y - a.rhs_mult_dt * P_dj[r].y;
        const double r0 = 1.0 / dps0, r1 = 1.0 / dps1;
        const double2 rd = make_double2(r0 * rx0, r1 * rx1);
        if (kStage) {
          const double vs10 = P_v1[r].x * r0, vs11 = P_v1[r].y * r1, vs20 = P_v2[r].x * r0, vs21 = P_v2[r].y * r1;
          u1 = make_double2((P_m11[r].x * vs10 + P_m12[r].x * vs20) * rx0, (P_m11[r].y * vs11 + P_m12[r].y * vs21) * rx1);
          u2 = make_double2((P_m21[r].x * vs10 + P_m22[r].x * vs20) * rx0, (P_m21[r].y * vs11 + P_m22[r].y * vs21) * rx1);
          cl = make_double2(P_sp[r].x * (dps0 - a.dt * P_dd[r].x), P_sp[r].y * (dps1 - a.dt * P_dd[r].y));
          rcl = make_double2(1.0 / cl.x, 1.0 / cl.y);
        }
        const int off = c * PCS + ppl * 16;
        if (cfg.U1 >= 0) *reinterpret_cast<double2*>(pp + cfg.U1 * PP_BYTES + off) = u1;
        if (cfg.U2 >= 0) *reinterpret_cast<double2*>(pp + cfg.U2 * PP_BYTES + off) = u2;
        if (cfg.CL >= 0) *reinterpret_cast<double2*>(pp + cfg.CL * PP_BYTES + off) = cl;
        if (cfg.RDP >= 0) *reinterpret_cast<double2*>(pp + cfg.RDP * PP_BYTES + off) = rd;
        if (cfg.RC >= 0) *reinterpret_cast<double2*>(pp + cfg.RC * PP_BYTES + off) = rcl;
      }
    }
  }

  // ---- per-thread DSS gather offsets (bytes inside an IN stage), two 16-bit offsets (8-byte units) per register ---------
  unsigned goff[NSLOT / 2];
  {
    TSE_UNROLL
    for (int s = 0; s < NSLOT; ++s) {
      const int code = evalid ? gsv[s] : -1;
      int off = ZERO_OFF;
      if (code >= 256) off = TILE_BYTES + ((qi * tb.hmax + (code - 256)) * KC + kk) * 8;
      else if (code >= 0) {
        const int p2 = qi * GPL + (code >> 4) * KC + kk, node = code & 15;
        off = p2 * 128 + ((((node >> 1) ^ swz(p2))) << 4) + (node & 1) * 8;
      }
      if (s & 1) goff[s >> 1] |= (unsigned)(off >> 3) << 16;
      else goff[s >> 1] = (unsigned)(off >> 3);
    }
  }
  auto gofs = [&](int s) -> int { return (int)(((s & 1) ? (goff[s >> 1] >> 16) : (goff[s >> 1] & 0xffffu)) << 3); };

  const int own_base = p * 128, own_sw = swz(p);
  // this warp's planes of tracer slot q2 are rows q2*GPL + wrow .. +BOX_ROWS-1 of the tile
  const int wrow = BOX_ROWS * (w % (GE / EPW)), wq0 = QW * (w / (GE / EPW));

  double sumc = 0.0;
  consumer_barrier();  // package visible to all consumer warps
  if (kStage) {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    TSE_UNROLL
    for (int c = 0; c < 8; c += 2) {
      const double2 x0 = lds128(pp + cfg.CL * PP_BYTES, c * PCS + pl * 16);
      const double2 x1 = lds128(pp + cfg.CL * PP_BYTES, (c + 1) * PCS + pl * 16);
      s0 += x0.x; s1 += x0.y; s2 += x1.x; s3 += x1.y;
    }
    sumc = (s0 + s1) + (s2 + s3);
  }
  const double cf = kS3 ? a.visc_coef * a.dp0[k] : 0.0;
  const double rkm1 = a.rkstage - 1.0, rrk = 1.0 / a.rkstage;

  double keep[16];  // STAGE3: cf*lap of the first item; TIME_AVG: Qdp(n0)
  const size_t pidx0 = (((size_t)g * NKC + kc) * Q + qi) * GPL + pl;
  for (int j = 0; j < nitems; ++j) {
    const int b = j % NST;
    mbar_wait(full_bar(b), (unsigned)((j / NST) & 1));
    const unsigned char* inb = smem + b * IN_STRIDE;
    const int it = j / NIN, which = j % NIN;
    const int q = it * QI + qi;
    const bool valid = evalid && q < Q;
    const size_t pidx = pidx0 + (size_t)it * QI * GPL;  // global plane index

    double S[16];
    TSE_UNROLL
    for (int c = 0; c < 8; ++c) {
      const double2 v = lds128(inb, own_base + ((c ^ own_sw) << 4));
      S[2 * c] = v.x;
      S[2 * c + 1] = v.y;
    }
#ifdef TSE_EXP_SKIP_GATHER  // timing experiment only: wrong results
    if (false) {
#else
    if (a.pending[which]) {  // DSS in the reference's unpack order: S, E, N, W edges, then SW, SE, NE, NW corners
#endif
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
    // limiter bounds of this plane (brought by the producer with the item: a global load here would sit on the critical path)
    double minp = 0.0, maxp = 0.0;
    if (kStage && which == NIN - 1) {
      minp = lds64(inb, BND_OFF + p * 8);
      maxp = lds64(inb, BND_OFF + (QI * GPL + p) * 8);
    }
    // Release the stage as soon as this thread's copies sit in registers (see "Stage release" at the top of this file).
#if TSE_RELEASE == 0
    fence_proxy_async_smem();
    mbar_arrive(empty_bar(b));
#elif TSE_RELEASE == 1
    {
      unsigned dep = 0;
      TSE_UNROLL
      for (int n = 0; n < 16; ++n) dep ^= (unsigned)__double2hiint(S[n]);
      if (kStage) dep ^= (unsigned)__double2hiint(minp) ^ (unsigned)__double2hiint(maxp);
      mbar_arrive_after(empty_bar(b), dep, (unsigned)a.zero);
    }
#else
    mbar_arrive(empty_bar(b));  // experiments only: unordered against the refill
#endif

    const bool last_of_iter = (which == NIN - 1);
    if (valid) {
      if (OP == OP_MINMAX || OP == OP_BIHARM_PRE) {
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) {
          const double2 rd = lds128(pp + cfg.RDP * PP_BYTES, c * PCS + pl * 16);
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
      } else if (OP == OP_MASS) {
        // J = sum_ij spheremp*Qdp of this plane (the per-element part of global_integral, global_norms_mod.F90:39-86)
        double J = 0.0;
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) {
          const double2 sp = lds128(elb + cfg.E1 * EL_BYTES, (c * GE + el) * 16);
          double2 rs = lds128(elb + cfg.RSPH * EL_BYTES, (c * GE + el) * 16);
          if (!a.pending[0]) rs = make_double2(1.0, 1.0);
          J = fma(sp.x, rs.x * S[2 * c], J);
          J = fma(sp.y, rs.y * S[2 * c + 1], J);
        }
        S[0] = J;
      } else if (OP == OP_TIME_AVG) {
        if (which == 0) {
          TSE_UNROLL
          for (int n = 0; n < 16; ++n) keep[n] = S[n];
        } else {
          TSE_UNROLL
          for (int c = 0; c < 8; ++c) {
            double2 rs = lds128(elb + cfg.RSPH * EL_BYTES, (c * GE + el) * 16);
            if (!a.pending[1]) rs = make_double2(1.0, 1.0);
            // the reference divides by rkstage (:657); multiplying by the rounded reciprocal differs by <= 1 ulp
            S[2 * c] = (keep[2 * c] + rkm1 * (rs.x * S[2 * c])) * rrk;
            S[2 * c + 1] = (keep[2 * c + 1] + rkm1 * (rs.y * S[2 * c + 1])) * rrk;
          }
        }
      } else if (kS3 && which == 0) {
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
#ifdef TSE_EXP_SKIP_S2CHECK  // timing experiment only: wrong results
        if (false) {
#else
        if (OP == OP_STAGE2) {
#endif
          // qmin = min(qmin, minval(Q)), qmax = max(qmax, maxval(Q)) with Q = rspheremp*DSS/dp (:779-792).  Away from tracer fronts
          // every Q already lies inside the bounds of stage 1: test that first (2 DSETP per node, no selects) and reduce only
          // where it fails -- the result is the same either way.
          double qv[16];
          TSE_UNROLL
          for (int c = 0; c < 8; ++c) {
            const double2 rd = lds128(pp + cfg.RDP * PP_BYTES, c * PCS + pl * 16);
            qv[2 * c] = S[2 * c] * rd.x;
            qv[2 * c + 1] = S[2 * c + 1] * rd.y;
          }
          if (any_outside(qv, minp, maxp)) {
            double mn0 = dmin(qv[0], qv[1]), mx0 = dmax(qv[0], qv[1]), mn1 = dmin(qv[2], qv[3]), mx1 = dmax(qv[2], qv[3]);
            TSE_UNROLL
            for (int n = 4; n < 16; n += 4) {
              mn0 = dmin(mn0, dmin(qv[n], qv[n + 1]));
              mx0 = dmax(mx0, dmax(qv[n], qv[n + 1]));
              mn1 = dmin(mn1, dmin(qv[n + 2], qv[n + 3]));
              mx1 = dmax(mx1, dmax(qv[n + 2], qv[n + 3]));
            }
            minp = dmin(minp, dmin(mn0, mn1));
            maxp = dmax(maxp, dmax(mx0, mx1));
          }
        }
        double y[16];
        asm volatile("" ::: "memory");
        flux_div(S, pp + cfg.U1 * PP_BYTES, pp + cfg.U2 * PP_BYTES, pl, D, y);
        asm volatile("" ::: "memory");
        const unsigned e1a = smem_u32 + (unsigned)(elb - smem) + (cfg.E1 < 0 ? 0 : cfg.E1) * EL_BYTES + el * 16;
        const unsigned e2a = smem_u32 + (unsigned)(elb - smem) + (cfg.E2 < 0 ? 0 : cfg.E2) * EL_BYTES + el * 16;
        TSE_UNROLL
        for (int c = 0; c < 8; ++c) {
          const double2 e1 = lds128v(e1a + c * GE * 16);
          const double2 e2 = lds128v(e2a + c * GE * 16);
          y[2 * c] = fma(-e2.x, y[2 * c], e1.x * S[2 * c]);
          y[2 * c + 1] = fma(-e2.y, y[2 * c + 1], e1.y * S[2 * c + 1]);
          if (kS3) {
            y[2 * c] += keep[2 * c];
            y[2 * c + 1] += keep[2 * c + 1];
          }
        }
        asm volatile("" ::: "memory");
#ifndef TSE_SKIP_LIMITER
        const unsigned cl_a = smem_u32 + (unsigned)(pp - smem) + (cfg.CL < 0 ? 0 : cfg.CL) * PP_BYTES + pl * 16;
        const unsigned rc_a = smem_u32 + (unsigned)(pp - smem) + (cfg.RC < 0 ? 0 : cfg.RC) * PP_BYTES + pl * 16;
        if (OP == OP_HYPERVIS) limiter2d_zero(y);
        else if (a.limiter8) limiter_y(y, cl_a, rc_a, sumc, minp, maxp);  // (:880: only option 8 limits inside euler_step)
#endif
        asm volatile("" ::: "memory");
        if (a.store_bounds) {  // the relaxed bounds (:1024-1029) are only read again by stage 2 (and by tse_get_qminmax)
          a.qmin[pidx] = minp;
          a.qmax[pidx] = maxp;
        }
        TSE_UNROLL
        for (int n = 0; n < 16; ++n) S[n] = y[n];
      }
    }

    if (OP == OP_MASS) {
      // two-limb fixed-point image of J (order-independent integer sums, the repro_sum idea, repro_sum_mod.F90:216-628) and the
      // running max of |J|; the 16 lanes that hold one tracer reduce first (EPW * KC = 16 planes)
      static_assert(OP != OP_MASS || EPW * KC == 16, "OP_MASS reduces over half-warps");
      const double J = valid ? S[0] : 0.0;
      const double x = scalbn(J, valid ? a.mass_shift[q] : 0);
      const double xi = trunc(x);
      long long hi = (long long)xi, lo = (long long)trunc(scalbn(x - xi, 40));
      unsigned long long mb = (unsigned long long)__double_as_longlong(fabs(J));
      TSE_UNROLL
      for (int o = 8; o > 0; o >>= 1) {
        hi += __shfl_xor_sync(0xffffffffu, hi, o);
        lo += __shfl_xor_sync(0xffffffffu, lo, o);
        const unsigned long long y = __shfl_xor_sync(0xffffffffu, mb, o);
        mb = y > mb ? y : mb;
      }
      if ((lane & 15) == 0 && q < Q) {
        // MASS_REP copies of the accumulators, picked by CTA: atomics on one address serialise in L2 (41 M of them on 105
        // addresses cost more than the whole pass); the copies are folded on the host (integer adds and max: order-free)
        const int rep = blockIdx.x % MASS_REP;
        atomicAdd(reinterpret_cast<unsigned long long*>(a.mass_acc + 2 * (rep * Q + q)), (unsigned long long)hi);
        atomicAdd(reinterpret_cast<unsigned long long*>(a.mass_acc + 2 * (rep * Q + q) + 1), (unsigned long long)lo);
        atomicMax(a.mass_maxbits + rep * Q + q, mb);
      }
    }

    if (kHasOut && last_of_iter) {
      // the store of this warp that last used this OUT buffer must have finished reading its rows
      unsigned char* const ob = outb + (it % NOUT) * TILE_BYTES;
      if (lane == 0) bulk_wait_read<NOUT - 1>();
      __syncwarp();
      TSE_UNROLL
      for (int c = 0; c < 8; ++c) *reinterpret_cast<double2*>(ob + own_base + ((c ^ own_sw) << 4)) = make_double2(S[2 * c], S[2 * c + 1]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int q0 = it * QI;
        for (int q2 = wq0; q2 < wq0 + QW && q0 + q2 < Q; ++q2)
          tma_store_2d(&maps.out, 0, (int)(row0 + (unsigned)(q0 + q2) * GPL) + wrow, smem_u32 + (unsigned)(ob - smem) + (q2 * GPL + wrow) * 128);
        bulk_commit();
      }
    }
  }
  if (kHasOut && lane == 0) bulk_wait0();  // (waiting for the reads only, wait_group.read, measured the same)

}

}  // namespace tse
