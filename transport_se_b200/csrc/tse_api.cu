// C ABI of the B200 tracer-advection path (see include/tse.h for the reference hooks each entry replaces).
//
// State lives in HBM for the whole run (tse_state).  Tracer fields use a pool of 4 buffers: the two Qdp time levels, the
// stage output (a stage cannot run in place because neighbours gather from its input) and the biharmonic temporary.
// A time-level slot may be "pending": its buffer holds the pre-DSS output of an RK stage and the DSS is applied by
// whichever kernel reads it next (or by an explicit resolve when the host asks for the field).
//
// Multi-GPU: one rank per GPU, elements split along the space-filling curve by the host.  The only data-path
// communication is the halo of each DSS (5 per tracer step, like the reference's 5 bndry_exchangeV calls): the boundary
// nodes facing another rank are packed into a send slab per neighbour rank and moved with ncclSend/ncclRecv into the ghost
// array the consumer kernels gather from.  Ghost values are bit copies and sums keep the reference's unpack order, so
// results are bit-for-bit identical for any number of GPUs.
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>

#include "tse.h"
#include "tse_kernels.cuh"
#include "tse_remap.cuh"
#include "tse_tile.cuh"
#include "tse_pipe.cuh"
#include "tse_dcmip.cuh"
#include "tse_diag.cuh"

using namespace tse;

namespace {

thread_local std::string g_err;

int fail(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

#define CU(call)                                                                                          \
  do {                                                                                                    \
    cudaError_t _e = (call);                                                                              \
    if (_e != cudaSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
  } while (0)
#define NC(call)                                                                                          \
  do {                                                                                                    \
    ncclResult_t _e = (call);                                                                             \
    if (_e != ncclSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #call, ncclGetErrorString(_e)); \
  } while (0)

const double kRearth = 6.376e6;  // physical_constants.F90:16-34
const double kRrearth = 1.0 / kRearth;

}  // namespace

struct tse_state {
  tse_config cfg{};
  int nelem = 0, ngroups = 0, npad = 0, Q = 0;
  cudaStream_t stream = nullptr;
  Geo geo{};
  Dvv dvv{};
  TileTables tiles{};
  NbrTables nbrt{};
  std::vector<int> h2i;  // host element -> internal element
  int* d_h2i = nullptr;
  int* d_gkey = nullptr;  // [internal element] global space-filling-curve index (host element index if none was given)
  // tracer buffers
  double* qbuf[4] = {nullptr, nullptr, nullptr, nullptr};
  double* qghost[4] = {nullptr, nullptr, nullptr, nullptr};  // halo of each buffer when it is pending (multi-GPU)
  CUtensorMap qmap[4];  // TMA view of each buffer: rows = planes of 16 doubles, box = BOX_ROWS planes, SWIZZLE_128B
  size_t qdoubles = 0;
  int slot_buf[3] = {-1, 0, 1};     // [tl] (1-based) -> buffer
  int slot_pending[3] = {0, 0, 0};  // [tl] buffer holds pre-DSS values
  // level fields
  size_t ldoubles = 0;
  double *vn0 = nullptr, *dp = nullptr, *divdp = nullptr, *divdp_proj = nullptr, *eta_dot = nullptr, *omega_p = nullptr, *lev_tmp = nullptr,
         *dp3d = nullptr, *ps_v = nullptr;
  double *qmin = nullptr, *qmax = nullptr, *qmin_loc = nullptr, *qmax_loc = nullptr;
  double *d_dp0 = nullptr, *d_dA = nullptr, *d_dB = nullptr;
  double hyai0_ps0 = 0, ps0 = 0;
  double* stage = nullptr;
  size_t stage_doubles = 0;
  // asynchronous upload of derived%vn0 / derived%dp (tse_set_derived): second copy of both fields, own staging area and stream, so
  // that the host->device copy of step n+1 runs under the kernels of step n
  double *vn0_alt = nullptr, *dp_alt = nullptr, *stage_up = nullptr;
  size_t stage_up_doubles = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_derived = nullptr, ev_alt_free = nullptr;
  int* d_err = nullptr;
  int* h_err = nullptr;             // pinned copy of d_err, refreshed after every vertical remap
  cudaEvent_t ev_err = nullptr;     // ... and the event that says the refresh has landed
  bool err_pending = false;
  long long launches = 0, stage_launches = 0, dev_bytes = 0;
  std::vector<void*> allocs;
  // halo exchange (multi-GPU)
  int nranks = 1, rank = 0;
  ncclComm_t comm = nullptr;
  struct Cycle { int peer, off, len, boff, blen; };  // ghost-slot range and bundle range exchanged with one neighbour rank
  std::vector<Cycle> cycles;
  int nghost = 0, nbundle = 0;
  int *d_send_src = nullptr, *d_mm_elem = nullptr;
  double *send_q = nullptr, *send_lev = nullptr, *ghost_lev = nullptr, *send_mm = nullptr, *ghost_mm = nullptr;
  long long halo_bytes = 0;  // bytes sent by this rank so far
  // overlap of the halo exchange with interior compute: groups owning a node that is sent come first, the exchange runs on
  // comm_stream while the remaining groups are processed
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_boundary = nullptr, ev_halo = nullptr;
  int *d_glist_b = nullptr, *d_glist_i = nullptr;
  int n_glist_b = 0, n_glist_i = 0;
  bool halo_outstanding = false;
  bool fused_step = false;  // inside tse_advec_tracers_remap_rk2
  // timers (lazily resolved CUDA events under the reference's GPTL names)
  std::map<std::string, double> timers;
  struct TimerRec { const char* name; cudaEvent_t a, b; };
  std::vector<TimerRec> timer_pending;
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t marks[16] = {};
  // prescribed-wind test case
  int test_case = 0;
  double *d_lon = nullptr, *d_lat = nullptr;
  DcmipTables dcmip{};
  std::vector<double> hv_hyai, hv_hybi, hv_hyam, hv_hybm, h_dA, h_dB;
  bool have_latlon = false;
  // diagnostics
  unsigned long long* d_maxbits = nullptr;
  long long* d_acc = nullptr;
  int* d_shift = nullptr;
  std::vector<int> mass_shift;  // binary scale of the fixed-point mass sum, per tracer (kept between calls)
};

namespace {

template <class T>
int dalloc(tse_state* s, T** p, size_t count) {
  void* v = nullptr;
  if (count == 0) count = 1;
  CU(cudaMalloc(&v, count * sizeof(T)));
  CU(cudaMemsetAsync(v, 0, count * sizeof(T), s->stream));
  s->allocs.push_back(v);
  s->dev_bytes += (long long)(count * sizeof(T));
  *p = (T*)v;
  return 0;
}
template <class T>
int upload(tse_state* s, T** p, const std::vector<T>& h) {
  if (dalloc(s, p, h.size())) return 1;
  if (!h.empty()) CU(cudaMemcpyAsync(*p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return 0;
}

inline int edge_node(int d, int t) {
  switch (d) {
    case SOUTH: return t;
    case EAST: return 3 + 4 * t;
    case NORTH: return 12 + t;
    default: return 4 * t;
  }
}
inline int corner_node(int d) { return d == SWEST ? 0 : d == SEAST ? 3 : d == NWEST ? 12 : 15; }

dim3 plane_grid(const tse_state* s) { return dim3((unsigned)(s->ngroups * NKC), (unsigned)((s->Q + QPB - 1) / QPB)); }

DssView view(const tse_state* s, int buf, int pending) {
  DssView v;
  v.q = s->qbuf[buf];
  v.ghost = s->qghost[buf];
  v.pending = pending;
  v.Q = s->Q;
  return v;
}

int pick_buffer(std::initializer_list<int> protect) {
  for (int b = 0; b < 4; ++b) {
    bool ok = true;
    for (int p : protect)
      if (p == b) ok = false;
    if (ok) return b;
  }
  return -1;
}

int check_tl(int tl) { return (tl == 1 || tl == 2) ? 0 : fail("time level %d out of range (1|2)", tl); }

// ---- timers -----------------------------------------------------------------------------------
cudaEvent_t get_event(tse_state* s) {
  if (!s->event_pool.empty()) {
    cudaEvent_t e = s->event_pool.back();
    s->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;  // timers degrade to no-ops (ScopedTimer checks)
  return e;
}
void resolve_timers(tse_state* s) {
  if (s->timer_pending.empty()) return;
  cudaStreamSynchronize(s->stream);
  for (auto& r : s->timer_pending) {
    float ms = 0;
    cudaEventElapsedTime(&ms, r.a, r.b);
    s->timers[r.name] += ms;
    s->event_pool.push_back(r.a);
    s->event_pool.push_back(r.b);
  }
  s->timer_pending.clear();
}
struct ScopedTimer {
  tse_state* s;
  tse_state::TimerRec r;
  ScopedTimer(tse_state* s_, const char* name) : s(s_) {
    if (s->timer_pending.size() > 8192) resolve_timers(s);
    r.name = name;
    r.a = get_event(s);
    r.b = get_event(s);
    if (r.a && r.b) cudaEventRecord(r.a, s->stream);
  }
  ~ScopedTimer() {
    if (!r.a || !r.b) {
      if (r.a) s->event_pool.push_back(r.a);
      if (r.b) s->event_pool.push_back(r.b);
      return;
    }
    cudaEventRecord(r.b, s->stream);
    s->timer_pending.push_back(r);
  }
};

// Negative layer thickness in vertical_remap (prim_advection_mod.F90:1319-1324: the reference calls abortmp).  The kernel raises
// d_err; a copy into pinned memory is queued right behind it.  Blocking entries (anything that hands results to the host, and
// tse_synchronize) report it after their own stream synchronisation; non-blocking entries report it as soon as the copy has
// landed, without waiting.  Reporting clears the flag, so the handle stays usable (the half-remapped field is the caller's
// problem, as after abortmp).
int report_remap_error(tse_state* s) {
  *s->h_err = 0;
  s->err_pending = false;
  CU(cudaMemsetAsync(s->d_err, 0, sizeof(int), s->stream));
  return fail("vertical_remap: negative layer thickness.  timestep or remap time too large");  // prim_advection_mod.F90:1323
}
int queue_error_readback(tse_state* s) {
  CU(cudaMemcpyAsync(s->h_err, s->d_err, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaEventRecord(s->ev_err, s->stream));
  s->err_pending = true;
  return 0;
}
// after the caller has synchronised s->stream
int check_device_error(tse_state* s) {
  if (!s->err_pending) return 0;
  CU(cudaEventSynchronize(s->ev_err));
  s->err_pending = false;
  return *s->h_err ? report_remap_error(s) : 0;
}
// without blocking
int poll_device_error(tse_state* s) {
  if (!s->err_pending || cudaEventQuery(s->ev_err) != cudaSuccess) return 0;
  s->err_pending = false;
  return *s->h_err ? report_remap_error(s) : 0;
}

// ---- halo exchange ----------------------------------------------------------------------------
// One ncclSend + ncclRecv per neighbour rank (a Cycle_t of the reference's schedule) and per payload, all payloads of one DSS in a
// single NCCL group (= one fused transfer kernel), like the reference packs all layers of a DSS into one message per neighbour.
struct Xfer {
  const double* send;
  double* recv;
  size_t unit;   // doubles per ghost slot (or per bundle)
  bool bundles;  // payload is indexed by (element, direction) bundle instead of by ghost slot
};
int exchange(tse_state* s, std::initializer_list<Xfer> xs) {
  if (s->cycles.empty()) return 0;
  if (!s->comm && std::getenv("TSE_PROFILE_NO_EXCHANGE")) {
    // profiling aid: time one rank's share of a multi-GPU run on a single GPU (ncu cannot wrap a multi-rank job).  The ghosts keep
    // whatever they hold, so the results are meaningless; every kernel, launch split and pack of the real run is there.
    CU(cudaEventRecord(s->ev_halo, s->comm_stream));
    s->halo_outstanding = true;
    return 0;
  }
  if (!s->comm) return fail("halo exchange: this rank has off-GPU neighbours but tse_comm_init was not called");
  NC(ncclGroupStart());
  ncclResult_t bad = ncclSuccess;
  for (const auto& c : s->cycles)
    for (const Xfer& x : xs) {
      const size_t off = (size_t)(x.bundles ? c.boff : c.off) * x.unit, cnt = (size_t)(x.bundles ? c.blen : c.len) * x.unit;
      ncclResult_t r = ncclSend(x.send + off, cnt, ncclDouble, c.peer, s->comm, s->comm_stream);
      if (r == ncclSuccess) r = ncclRecv(x.recv + off, cnt, ncclDouble, c.peer, s->comm, s->comm_stream);
      if (r != ncclSuccess && bad == ncclSuccess) bad = r;
      s->halo_bytes += (long long)cnt * 8;
    }
  {
    const ncclResult_t r = ncclGroupEnd();  // always closed, also after a failed send/recv
    if (bad == ncclSuccess) bad = r;
  }
  if (bad != ncclSuccess) return fail("halo exchange: %s", ncclGetErrorString(bad));
  CU(cudaEventRecord(s->ev_halo, s->comm_stream));
  s->halo_outstanding = true;
  return 0;
}
// the compute stream must not read ghosts before the exchange has landed
int wait_halo(tse_state* s) {
  if (!s->halo_outstanding) return 0;
  ScopedTimer tm(s, "bndry_exchange");  // time the compute stream is stalled on the halo (what is not hidden by interior compute)
  CU(cudaStreamWaitEvent(s->stream, s->ev_halo, 0));
  s->halo_outstanding = false;
  return 0;
}
// the comm stream starts packing once the boundary groups of the producing kernel are done
int comm_after_boundary(tse_state* s) {
  CU(cudaEventRecord(s->ev_boundary, s->stream));
  CU(cudaStreamWaitEvent(s->comm_stream, s->ev_boundary, 0));
  return 0;
}
int pack_tracer(tse_state* s, int buf) {
  const size_t total = (size_t)s->nghost * s->Q * NLEV;
  k_pack_tracer<<<(unsigned)((total + 255) / 256), 256, 0, s->comm_stream>>>(s->qbuf[buf], s->d_send_src, s->nghost, s->Q, s->send_q);
  ++s->launches;
  CU(cudaGetLastError());
  return 0;
}
int pack_minmax(tse_state* s) {
  const size_t total = (size_t)s->nbundle * s->Q * NLEV;
  k_pack_minmax<<<(unsigned)((total + 255) / 256), 256, 0, s->comm_stream>>>(s->qmin_loc, s->qmax_loc, s->d_mm_elem, s->nbundle, s->Q, s->send_mm);
  ++s->launches;
  CU(cudaGetLastError());
  return 0;
}
int pack_level(tse_state* s, const double* f) {
  const size_t total = (size_t)s->nghost * NLEV;
  k_pack_level<<<(unsigned)((total + 255) / 256), 256, 0, s->comm_stream>>>(f, s->geo.spheremp, s->d_send_src, s->nghost, s->send_lev);
  ++s->launches;
  CU(cudaGetLastError());
  return 0;
}
Xfer xfer_tracer(tse_state* s, int buf) { return Xfer{s->send_q, s->qghost[buf], (size_t)s->Q * NLEV, false}; }
Xfer xfer_minmax(tse_state* s) { return Xfer{s->send_mm, s->ghost_mm, (size_t)2 * s->Q * NLEV, true}; }
Xfer xfer_level(tse_state* s) { return Xfer{s->send_lev, s->ghost_lev, (size_t)NLEV, false}; }

TileArgs tile_args(const tse_state* s) {
  TileArgs a{};
  a.vn0 = s->vn0; a.dp = s->dp; a.divdp = s->divdp; a.divdp_proj = s->divdp_proj;
  a.dp0 = s->d_dp0;
  a.qmin = s->qmin; a.qmax = s->qmax; a.qmin_loc = s->qmin_loc; a.qmax_loc = s->qmax_loc;
  a.Q = s->Q;
  a.rkstage = 3.0;
  a.store_bounds = 1;
  a.limiter8 = s->cfg.limiter_option == 8 ? 1 : 0;
  return a;
}
void set_src(const tse_state* s, TileArgs& a, int i, int buf, int pending) {
  a.src[i] = s->qbuf[buf];
  a.ghost[i] = s->qghost[buf];
  a.pending[i] = pending;
}
const CUtensorMap& map_of(const tse_state* s, const double* buf) {
  for (int b = 0; b < 4; ++b)
    if (s->qbuf[b] == buf) return s->qmap[b];
  return s->qmap[0];
}
template <int OP>
void launch_tile(tse_state* s, TileArgs a, const int* glist = nullptr, int ngl = 0) {
  a.glist = glist;
  const int ng = glist ? ngl : s->ngroups;
  if (ng == 0) return;
  PipeMaps pm;
  pm.in[0] = map_of(s, a.src[0]);
  pm.in[1] = map_of(s, a.src[1] ? a.src[1] : a.src[0]);
  pm.out = map_of(s, a.out ? a.out : a.src[0]);
  k_pipe<OP><<<ng * NKC, pipe_threads(OP), pipe_smem_bytes(OP, s->tiles.hmax), s->stream>>>(pm, s->geo, s->dvv, s->tiles, a);
  ++s->launches;
}
// producer launch of a field whose boundary nodes are exchanged: boundary groups, then (after `start_comm` queued the pack and
// the exchange on the comm stream) the interior groups
template <int OP, class F>
int launch_tile_overlapped(tse_state* s, const TileArgs& a, F start_comm) {
  if (s->cycles.empty()) {
    launch_tile<OP>(s, a);
    CU(cudaGetLastError());
    return 0;
  }
  launch_tile<OP>(s, a, s->d_glist_b, s->n_glist_b);
  CU(cudaGetLastError());
  if (comm_after_boundary(s)) return 1;
  if (start_comm()) return 1;
  launch_tile<OP>(s, a, s->d_glist_i, s->n_glist_i);
  CU(cudaGetLastError());
  return 0;
}
// neighbor_minmax (viscosity_mod.F90:748-816) after the element extrema of off-GPU neighbours have arrived: 9-way min/max
int neighbor_minmax(tse_state* s) {
  if (wait_halo(s)) return 1;
  k_nbr_minmax<<<s->ngroups * NKC, NBQ * GPL, nbr_smem_bytes(s->nbrt.xmax), s->stream>>>(s->geo, s->nbrt, s->Q, s->qmin_loc, s->qmax_loc,
                                                                                       s->ghost_mm, s->qmin, s->qmax);
  ++s->launches;
  CU(cudaGetLastError());
  return 0;
}

int resolve_slot(tse_state* s, int tl) {
  if (!s->slot_pending[tl]) return 0;
  if (wait_halo(s)) return 1;
  const int other = s->slot_buf[3 - tl], in = s->slot_buf[tl];
  const int out = pick_buffer({in, other});
  TileArgs a = tile_args(s);
  set_src(s, a, 0, in, 1);
  a.out = s->qbuf[out];
  launch_tile<OP_RESOLVE>(s, a);
  CU(cudaGetLastError());
  s->slot_buf[tl] = out;
  s->slot_pending[tl] = 0;
  return 0;
}

// DSS of the extra level field of euler_step (prim_advection_mod.F90:913-919, 943-958)
double** level_field(tse_state* s, int DSSopt) {
  return DSSopt == TSE_DSS_ETA ? &s->eta_dot : DSSopt == TSE_DSS_OMEGA ? &s->omega_p : DSSopt == TSE_DSS_DIV_VDP_AVE ? &s->divdp_proj : nullptr;
}
int dss_level_field(tse_state* s, int DSSopt) {  // the halo of the field (ghost_lev) must already have been exchanged
  double** f = DSSopt == TSE_DSS_ETA ? &s->eta_dot : DSSopt == TSE_DSS_OMEGA ? &s->omega_p : DSSopt == TSE_DSS_DIV_VDP_AVE ? &s->divdp_proj : nullptr;
  if (DSSopt != TSE_DSS_NO_VAR && !f) return fail("tse_euler_step: DSSopt=%d", DSSopt);
  if (!f) return 0;
  if (wait_halo(s)) return 1;
  k_dss_level<<<s->ngroups, DSL_THREADS, dss_level_smem_bytes(s->tiles.hmax), s->stream>>>(s->geo, s->tiles, *f, s->ghost_lev, s->lev_tmp);
  ++s->launches;
  CU(cudaGetLastError());
  std::swap(*f, s->lev_tmp);
  return 0;
}

}  // namespace

extern "C" {

const char* tse_last_error(void) { return g_err.c_str(); }

int tse_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int tse_init(const tse_config* cfg, const tse_geometry* geom, const tse_connectivity* conn, const tse_hvcoord* hv, const double* dvv,
             tse_handle* out) {
  if (!cfg || !geom || !conn || !hv || !dvv || !out) return fail("tse_init: null argument");
  if (cfg->np != NP || cfg->nlev != NLEV) return fail("tse_init: built for np=4, nlev=72 (got np=%d nlev=%d)", cfg->np, cfg->nlev);
  // limiter_option: 8 = limiter_optim_iter_full; any other value advects without a limiter, exactly like the reference's CPU
  // path (prim_advection_mod.F90:858,880 test for 8 only; limiter2d_zero / limiter2d_minmax are never called there).
  // hypervis_subcycle_q is read, broadcast and printed by the reference but used nowhere on its CPU path; the one rule is kept:
  if (cfg->limiter_option == 8 && cfg->hypervis_subcycle_q != 1)
    return fail("tse_init: limiter 8 requires hypervis_subcycle_q=1 (namelist_mod.F90:688-692)");
  if (cfg->vert_remap_q_alg == 2) return fail("tse_init: vert_remap_q_alg=2 is not implemented");
  if (cfg->qsize < 1 || cfg->qsize > cfg->qsize_d) return fail("tse_init: qsize=%d qsize_d=%d", cfg->qsize, cfg->qsize_d);
  if (cfg->nelemd < 1) return fail("tse_init: nelemd=%d", cfg->nelemd);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail("tse_init: no CUDA device (this library has no CPU path)");
  if (cfg->device >= 0) CU(cudaSetDevice(cfg->device));

  // every early return below releases what has been created so far (streams, events, HBM: tens of GB at ne120)
  struct Guard {
    tse_state* p;
    ~Guard() { if (p) tse_finalize(p); }
  } guard{new tse_state};
  tse_state* s = guard.p;
  s->cfg = *cfg;
  s->nelem = cfg->nelemd;
  s->ngroups = (s->nelem + GE - 1) / GE;
  s->npad = s->ngroups * GE;
  s->Q = cfg->qsize;
  CU(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  std::memcpy(s->dvv.d, dvv, sizeof s->dvv.d);
  const int ne = s->nelem;

  // internal element order: along the space-filling curve when the host provides it
  std::vector<int> order(ne);
  std::iota(order.begin(), order.end(), 0);
  if (conn->sfc_index) std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return conn->sfc_index[a] < conn->sfc_index[b]; });
  s->h2i.assign(ne, 0);
  for (int i = 0; i < ne; ++i) s->h2i[order[i]] = i;
  if (upload(s, &s->d_h2i, s->h2i)) return 1;
  {
    std::vector<int> gkey(s->npad, 0);
    for (int eh = 0; eh < ne; ++eh) gkey[s->h2i[eh]] = conn->sfc_index ? conn->sfc_index[eh] : eh;
    if (upload(s, &s->d_gkey, gkey)) return 1;
  }

  // geometry
  const size_t n16 = (size_t)s->npad * 16;
  std::vector<double> sp(n16, 1.0), rsp(n16, 1.0), rmr(n16, 0.0), mD((size_t)s->npad * 64, 0.0), T((size_t)s->npad * 48, 0.0);
  for (int eh = 0; eh < ne; ++eh) {
    const int e = s->h2i[eh];
    for (int n = 0; n < 16; ++n) {
      const size_t hi = (size_t)eh * 16 + n, di = (size_t)e * 16 + n;
      sp[di] = geom->spheremp[hi];
      rsp[di] = geom->rspheremp[hi];
      rmr[di] = geom->rmetdet[hi] * kRrearth;
      const double* Di = geom->Dinv + hi * 4;  // Dinv(a,b) at [a + 2b]
      const double d11 = Di[0], d21 = Di[1], d12 = Di[2], d22 = Di[3];
      const double md = geom->metdet[hi];
      mD[(size_t)e * 64 + n] = md * d11;
      mD[(size_t)e * 64 + 16 + n] = md * d12;
      mD[(size_t)e * 64 + 32 + n] = md * d21;
      mD[(size_t)e * 64 + 48 + n] = md * d22;
      const double f = geom->spheremp[hi] * kRrearth * kRrearth;
      T[(size_t)e * 48 + n] = f * (d11 * d11 + d12 * d12);
      T[(size_t)e * 48 + 16 + n] = f * (d11 * d21 + d12 * d22);
      T[(size_t)e * 48 + 32 + n] = f * (d21 * d21 + d22 * d22);
    }
  }
  double *d_sp, *d_rsp, *d_rmr, *d_mD, *d_T;
  if (upload(s, &d_sp, sp) || upload(s, &d_rsp, rsp) || upload(s, &d_rmr, rmr) || upload(s, &d_mD, mD) || upload(s, &d_T, T)) return 1;

  // ---- DSS gather table from the reference's put/get maps (edge_mod.F90:366-511 pack, :648-742 unpack) ----
  // A slot of the edge buffer inside a cycle range [ptrP, ptrP+lengthP) belongs to a neighbour rank: what this rank packs
  // there is sent (send_src), what it unpacks from there is the neighbour's value (ghost).
  const int nbuf = conn->nbuf;
  std::vector<int> slot_src(nbuf, -1), ghost_id(nbuf, -1);
  int nghost = 0;
  for (int c = 0; c < conn->ncycles; ++c) {
    if (conn->cyc_ptr[c] < 0 || conn->cyc_ptr[c] + conn->cyc_len[c] > nbuf) return fail("tse_init: exchange cycle out of range");
    for (int i = 0; i < conn->cyc_len[c]; ++i) ghost_id[conn->cyc_ptr[c] + i] = nghost++;
  }
  s->nghost = nghost;
  std::vector<int> send_src(nghost, 0);
  for (int eh = 0; eh < ne; ++eh)
    for (int d = 0; d < 8; ++d) {
      const int pm = conn->putmapP[eh * 8 + d];
      if (pm < 0) continue;
      const int len = d < 4 ? NP : 1;
      if (pm + len > nbuf) return fail("tse_init: putmapP out of range");
      for (int i = 0; i < len; ++i) {
        const int t = (d < 4) ? (conn->reverse[eh * 8 + d] ? NP - 1 - i : i) : 0;
        const int code = (s->h2i[eh] << 4) | (d < 4 ? edge_node(d, t) : corner_node(d));
        if (ghost_id[pm + i] >= 0) send_src[ghost_id[pm + i]] = code;
        else slot_src[pm + i] = code;
      }
    }
  // bundles = (element, direction) pairs facing another rank, numbered in ghost-slot order (identical on both sides)
  std::vector<std::pair<int, int>> bundle_key;  // (first ghost id, host element)
  for (int eh = 0; eh < ne; ++eh)
    for (int d = 0; d < 8; ++d) {
      const int gm = conn->getmapP[eh * 8 + d];
      if (gm >= 0 && gm < nbuf && ghost_id[gm] >= 0) bundle_key.push_back({ghost_id[gm], eh});
    }
  std::sort(bundle_key.begin(), bundle_key.end());
  s->nbundle = (int)bundle_key.size();
  std::map<int, int> bundle_of;  // first ghost id -> bundle index
  std::vector<int> mm_elem(s->nbundle, 0);
  for (int b = 0; b < s->nbundle; ++b) {
    bundle_of[bundle_key[b].first] = b;
    mm_elem[b] = s->h2i[bundle_key[b].second];
  }
  {
    int goff = 0;
    for (int c = 0; c < conn->ncycles; ++c) {
      tse_state::Cycle cy;
      cy.peer = conn->cyc_rank[c];
      cy.off = goff;
      cy.len = conn->cyc_len[c];
      goff += cy.len;
      auto lo = std::lower_bound(bundle_key.begin(), bundle_key.end(), std::make_pair(cy.off, -1));
      auto hi = std::lower_bound(bundle_key.begin(), bundle_key.end(), std::make_pair(cy.off + cy.len, -1));
      cy.boff = (int)(lo - bundle_key.begin());
      cy.blen = (int)(hi - lo);
      s->cycles.push_back(cy);
    }
  }
  std::vector<int> gsrc((size_t)s->npad * NSLOT, -1), nbr8((size_t)s->npad * 8, -1);
  const int unpack_edges[4] = {SOUTH, EAST, NORTH, WEST};
  const int unpack_corners[4] = {SWEST, SEAST, NEAST, NWEST};
  for (int eh = 0; eh < ne; ++eh) {
    const int e = s->h2i[eh];
    auto src_of = [&](int b) -> int {
      if (b < 0 || b >= nbuf) return -1;
      if (ghost_id[b] >= 0) return -(ghost_id[b] + 2);
      return slot_src[b];
    };
    auto nbr_of = [&](int gm) -> int {  // neighbour element for min/max: local element, ghost bundle or none
      const int b = src_of(gm);
      if (b >= 0) return b >> 4;
      if (b == -1) return -1;
      return -(bundle_of[ghost_id[gm]] + 2);
    };
    for (int x = 0; x < 4; ++x) {
      const int gm = conn->getmapP[eh * 8 + unpack_edges[x]];
      for (int i = 0; i < 4; ++i) gsrc[(size_t)e * NSLOT + 4 * x + i] = gm < 0 ? -1 : src_of(gm + i);
      nbr8[(size_t)e * 8 + x] = gm < 0 ? -1 : nbr_of(gm);
    }
    for (int x = 0; x < 4; ++x) {
      const int gm = conn->getmapP[eh * 8 + unpack_corners[x]];
      gsrc[(size_t)e * NSLOT + 16 + x] = gm < 0 ? -1 : src_of(gm);
      nbr8[(size_t)e * 8 + 4 + x] = gm < 0 ? -1 : nbr_of(gm);
    }
  }
  if (nghost > 0) {
    std::vector<char> is_b(s->ngroups, 0);
    for (int i = 0; i < nghost; ++i) is_b[(send_src[i] >> 4) / GE] = 1;
    for (int b = 0; b < s->nbundle; ++b) is_b[mm_elem[b] / GE] = 1;
    std::vector<int> gb, gi;
    for (int g = 0; g < s->ngroups; ++g) (is_b[g] ? gb : gi).push_back(g);
    s->n_glist_b = (int)gb.size();
    s->n_glist_i = (int)gi.size();
    if (upload(s, &s->d_glist_b, gb) || upload(s, &s->d_glist_i, gi)) return 1;
    {
      // highest priority: the pack and NCCL kernels must get SM slots as interior CTAs retire; at equal priority they sit
      // behind the whole interior grid and the exchange is not overlapped at all (measured: 0.6 ms exposed per DSS at ne120 / 8 GPUs)
      int lo = 0, hi = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CU(cudaStreamCreateWithPriority(&s->comm_stream, cudaStreamNonBlocking, hi));
    }
    CU(cudaEventCreateWithFlags(&s->ev_boundary, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s->ev_halo, cudaEventDisableTiming));
  }
  int *d_gsrc, *d_nbr8;
  if (upload(s, &d_gsrc, gsrc) || upload(s, &d_nbr8, nbr8) || upload(s, &s->d_send_src, send_src) || upload(s, &s->d_mm_elem, mm_elem))
    return 1;
  {
    // group-local view of the gather table: sources inside the 16-element group are read from the shared-memory tile,
    // everything else (other groups, other GPUs) goes through a per-group halo list
    std::vector<int> gsrc_t((size_t)s->npad * NSLOT, -1), halo_off(s->ngroups + 1, 0), halo_src;
    int hmax = 1;
    for (int g = 0; g < s->ngroups; ++g) {
      std::vector<int> ext;
      for (int el = 0; el < GE; ++el)
        for (int x = 0; x < NSLOT; ++x) {
          const int code = gsrc[((size_t)g * GE + el) * NSLOT + x];
          if (code == -1) continue;
          if (code >= 0 && (code >> 4) / GE == g) continue;
          ext.push_back(code);
        }
      std::sort(ext.begin(), ext.end());
      ext.erase(std::unique(ext.begin(), ext.end()), ext.end());
      for (int el = 0; el < GE; ++el)
        for (int x = 0; x < NSLOT; ++x) {
          const size_t i = ((size_t)g * GE + el) * NSLOT + x;
          const int code = gsrc[i];
          if (code == -1) continue;
          if (code >= 0 && (code >> 4) / GE == g) gsrc_t[i] = (((code >> 4) % GE) << 4) | (code & 15);
          else gsrc_t[i] = 256 + (int)(std::lower_bound(ext.begin(), ext.end(), code) - ext.begin());
        }
      halo_off[g + 1] = halo_off[g] + (int)ext.size();
      halo_src.insert(halo_src.end(), ext.begin(), ext.end());
      hmax = std::max(hmax, (int)ext.size());
    }
    if (halo_src.empty()) halo_src.push_back(-1);
    int *d_a, *d_b, *d_c;
    if (upload(s, &d_a, gsrc_t) || upload(s, &d_b, halo_off) || upload(s, &d_c, halo_src)) return 1;
    s->tiles.gsrc_t = d_a; s->tiles.halo_off = d_b; s->tiles.halo_src = d_c; s->tiles.hmax = hmax;
    if (pipe_smem_bytes(OP_STAGE2, hmax) > 227 * 1024 || pipe_smem_bytes(OP_STAGE3, hmax) > 227 * 1024)
      return fail("tse_init: halo of %d nodes per group does not fit in shared memory", hmax);
    if (tile_in_bytes(hmax) / 8 >= 65536) return fail("tse_init: halo of %d nodes per group overflows the gather offsets", hmax);
  }
  {
    // group-local view of the neighbour-element table for k_nbr_minmax: neighbours inside the group are read from the group's
    // own tile in shared memory, the others (other groups, ghost bundles of other GPUs) through a per-group list
    std::vector<int> nbr_t((size_t)s->npad * 8, -1), ext_off(s->ngroups + 1, 0), ext_src;
    int xmax = 1;
    for (int g = 0; g < s->ngroups; ++g) {
      std::vector<int> ext;
      for (int el = 0; el < GE; ++el)
        for (int d = 0; d < 8; ++d) {
          const int b = nbr8[((size_t)g * GE + el) * 8 + d];
          if (b == -1 || (b >= 0 && b / GE == g)) continue;
          ext.push_back(b);
        }
      std::sort(ext.begin(), ext.end());
      ext.erase(std::unique(ext.begin(), ext.end()), ext.end());
      for (int el = 0; el < GE; ++el)
        for (int d = 0; d < 8; ++d) {
          const size_t i = ((size_t)g * GE + el) * 8 + d;
          const int b = nbr8[i];
          if (b == -1) continue;
          if (b >= 0 && b / GE == g) nbr_t[i] = b % GE;
          else nbr_t[i] = 256 + (int)(std::lower_bound(ext.begin(), ext.end(), b) - ext.begin());
        }
      ext_off[g + 1] = ext_off[g] + (int)ext.size();
      ext_src.insert(ext_src.end(), ext.begin(), ext.end());
      xmax = std::max(xmax, (int)ext.size());
    }
    if (ext_src.empty()) ext_src.push_back(-1);
    int *d_a, *d_b, *d_c;
    if (upload(s, &d_a, nbr_t) || upload(s, &d_b, ext_off) || upload(s, &d_c, ext_src)) return 1;
    s->nbrt.nbr_t = d_a; s->nbrt.ext_off = d_b; s->nbrt.ext_src = d_c; s->nbrt.xmax = xmax;
    if (nbr_smem_bytes(xmax) > 200 * 1024) return fail("tse_init: %d external neighbour elements per group do not fit in shared memory", xmax);
    CU(cudaFuncSetAttribute(k_nbr_minmax, cudaFuncAttributeMaxDynamicSharedMemorySize, nbr_smem_bytes(xmax)));
  }
  s->geo.spheremp = d_sp; s->geo.rspheremp = d_rsp; s->geo.rmr = d_rmr; s->geo.mD = d_mD; s->geo.T = d_T;
  s->geo.gsrc = d_gsrc; s->geo.nbr8 = d_nbr8; s->geo.nelem = ne; s->geo.ngroups = s->ngroups;

  // vertical coordinate
  std::vector<double> dp0(NLEV), dA(NLEV), dB(NLEV);
  for (int k = 0; k < NLEV; ++k) {
    dA[k] = (hv->hyai[k + 1] - hv->hyai[k]) * hv->ps0;
    dB[k] = hv->hybi[k + 1] - hv->hybi[k];
    dp0[k] = (hv->hyai[k + 1] - hv->hyai[k]) * hv->ps0 + (hv->hybi[k + 1] - hv->hybi[k]) * hv->ps0;  // prim_advection_mod.F90:818-820
  }
  s->h_dA = dA;
  s->h_dB = dB;
  s->hyai0_ps0 = hv->hyai[0] * hv->ps0;
  s->ps0 = hv->ps0;
  s->hv_hyai.assign(hv->hyai, hv->hyai + NLEV + 1);
  s->hv_hybi.assign(hv->hybi, hv->hybi + NLEV + 1);
  if (hv->hyam && hv->hybm) {
    s->hv_hyam.assign(hv->hyam, hv->hyam + NLEV);
    s->hv_hybm.assign(hv->hybm, hv->hybm + NLEV);
  }
  if (geom->lat && geom->lon) {
    std::vector<double> la(n16, 0.0), lo(n16, 0.0);
    for (int eh = 0; eh < ne; ++eh)
      for (int n = 0; n < 16; ++n) {
        la[(size_t)s->h2i[eh] * 16 + n] = geom->lat[(size_t)eh * 16 + n];
        lo[(size_t)s->h2i[eh] * 16 + n] = geom->lon[(size_t)eh * 16 + n];
      }
    if (upload(s, &s->d_lat, la) || upload(s, &s->d_lon, lo)) return 1;
    s->have_latlon = true;
  }
  if (upload(s, &s->d_dp0, dp0) || upload(s, &s->d_dA, dA) || upload(s, &s->d_dB, dB)) return 1;
  s->mass_shift.assign(s->Q, 0);
  if (dalloc(s, &s->d_maxbits, (size_t)MASS_REP * s->Q) || dalloc(s, &s->d_acc, (size_t)2 * MASS_REP * s->Q) || dalloc(s, &s->d_shift, (size_t)s->Q)) return 1;

  // state
  s->ldoubles = (size_t)s->ngroups * NKC * GPL * 16;
  s->qdoubles = s->ldoubles * s->Q;
  if (s->qdoubles >= ((size_t)1 << 35)) return fail("tse_init: tracer field too large for one GPU (%zu doubles)", s->qdoubles);
  for (int b = 0; b < 4; ++b)
    if (dalloc(s, &s->qbuf[b], s->qdoubles)) return 1;
  {
    // tensor maps for the TMA tile loads/stores of k_pipe (driver entry point resolved through the runtime: no -lcuda)
    typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail("tse_init: cuTensorMapEncodeTiled is not available in this driver");
    const cuuint64_t gdim[2] = {16, (cuuint64_t)(s->qdoubles / 16)};
    const cuuint64_t gstride[1] = {128};
    const cuuint32_t box[2] = {16, (cuuint32_t)BOX_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    for (int b = 0; b < 4; ++b) {
      const CUresult r = ((EncodeTiled)fn)(&s->qmap[b], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, s->qbuf[b], gdim, gstride, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail("tse_init: cuTensorMapEncodeTiled failed (%d)", (int)r);
    }
  }
  if (dalloc(s, &s->vn0, 2 * s->ldoubles) || dalloc(s, &s->dp, s->ldoubles) || dalloc(s, &s->divdp, s->ldoubles) ||
      dalloc(s, &s->divdp_proj, s->ldoubles) || dalloc(s, &s->eta_dot, s->ldoubles) || dalloc(s, &s->omega_p, s->ldoubles) ||
      dalloc(s, &s->lev_tmp, s->ldoubles) || dalloc(s, &s->dp3d, s->ldoubles) || dalloc(s, &s->ps_v, (size_t)s->npad * 16))
    return 1;
  const size_t nplanes = s->qdoubles / 16;
  if (dalloc(s, &s->qmin, nplanes) || dalloc(s, &s->qmax, nplanes) || dalloc(s, &s->qmin_loc, nplanes) || dalloc(s, &s->qmax_loc, nplanes))
    return 1;
  if (dalloc(s, &s->d_err, 1)) return 1;
  CU(cudaMallocHost(&s->h_err, sizeof(int)));
  *s->h_err = 0;
  CU(cudaEventCreateWithFlags(&s->ev_err, cudaEventDisableTiming));
  if (nghost > 0) {
    const size_t gq = (size_t)nghost * s->Q * NLEV;
    for (int b = 0; b < 4; ++b)
      if (dalloc(s, &s->qghost[b], gq)) return 1;
    if (dalloc(s, &s->send_q, gq) || dalloc(s, &s->send_lev, (size_t)nghost * NLEV) || dalloc(s, &s->ghost_lev, (size_t)nghost * NLEV) ||
        dalloc(s, &s->send_mm, (size_t)s->nbundle * 2 * s->Q * NLEV) || dalloc(s, &s->ghost_mm, (size_t)s->nbundle * 2 * s->Q * NLEV))
      return 1;
  }
  {  // staging buffer for host<->device layout conversion: whole elements, ~256 MB at most
    const size_t per_elem = (size_t)16 * NLEV * std::max(s->Q, 2) + 16;
    const size_t ne_chunk = std::max<size_t>(1, std::min<size_t>(ne, ((size_t)32 << 20) / per_elem));
    s->stage_doubles = ne_chunk * per_elem;
    if (dalloc(s, &s->stage, s->stage_doubles)) return 1;
  }
  {
    const size_t per_elem = (size_t)16 * NLEV * 2;
    const size_t ne_chunk = std::max<size_t>(1, std::min<size_t>(ne, ((size_t)16 << 20) / per_elem));
    s->stage_up_doubles = 2 * ne_chunk * per_elem;  // two halves of 128 MB: the copy of one chunk overlaps the relayout of the previous
    if (dalloc(s, &s->stage_up, s->stage_up_doubles) || dalloc(s, &s->vn0_alt, 2 * s->ldoubles) || dalloc(s, &s->dp_alt, s->ldoubles)) return 1;
    {
      // highest priority: the small relayout kernels must not queue behind the grids of the step that is running
      int lo = 0, hi = 0;
      CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CU(cudaStreamCreateWithPriority(&s->copy_stream, cudaStreamNonBlocking, hi));
    }
    CU(cudaEventCreateWithFlags(&s->ev_derived, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&s->ev_alt_free, cudaEventDisableTiming));
  }
  CU(cudaFuncSetAttribute(k_vertical_remap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RM_SMEM));
  const int hm = s->tiles.hmax;
#define TSE_TILE_SMEM(OP) CU(cudaFuncSetAttribute(k_pipe<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, pipe_smem_bytes(OP, hm)))
  TSE_TILE_SMEM(OP_MINMAX);
  TSE_TILE_SMEM(OP_STAGE1);
  TSE_TILE_SMEM(OP_STAGE2);
  TSE_TILE_SMEM(OP_STAGE3);
  TSE_TILE_SMEM(OP_BIHARM_PRE);
  TSE_TILE_SMEM(OP_TIME_AVG);
  TSE_TILE_SMEM(OP_RESOLVE);
  TSE_TILE_SMEM(OP_MASS);
  TSE_TILE_SMEM(OP_HYPERVIS);
  CU(cudaFuncSetAttribute(k_dss_level, cudaFuncAttributeMaxDynamicSharedMemorySize, dss_level_smem_bytes(hm)));
#undef TSE_TILE_SMEM
  CU(cudaStreamSynchronize(s->stream));
  guard.p = nullptr;
  *out = s;
  return 0;
}

int tse_finalize(tse_handle s) {
  if (!s) return 0;
  if (s->stream) cudaStreamSynchronize(s->stream);
  resolve_timers(s);
  if (s->comm_stream) cudaStreamSynchronize(s->comm_stream);
  if (s->comm) ncclCommDestroy(s->comm);
  if (s->ev_boundary) cudaEventDestroy(s->ev_boundary);
  if (s->ev_halo) cudaEventDestroy(s->ev_halo);
  if (s->comm_stream) cudaStreamDestroy(s->comm_stream);
  if (s->copy_stream) {
    cudaStreamSynchronize(s->copy_stream);
    cudaStreamDestroy(s->copy_stream);
  }
  if (s->ev_derived) cudaEventDestroy(s->ev_derived);
  if (s->ev_alt_free) cudaEventDestroy(s->ev_alt_free);
  for (cudaEvent_t e : s->event_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : s->marks)
    if (e) cudaEventDestroy(e);
  for (void* p : s->allocs) cudaFree(p);
  if (s->h_err) cudaFreeHost(s->h_err);
  if (s->ev_err) cudaEventDestroy(s->ev_err);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return 0;
}

int tse_synchronize(tse_handle s) {
  if (!s) return fail("tse_synchronize: null handle");
  if (s->copy_stream) CU(cudaStreamSynchronize(s->copy_stream));
  if (s->comm_stream) CU(cudaStreamSynchronize(s->comm_stream));
  CU(cudaStreamSynchronize(s->stream));
  return check_device_error(s);
}

int tse_comm_unique_id(void* id128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NC(ncclGetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)));
  return 0;
}

int tse_comm_init(tse_handle s, int nranks, int rank, const void* id128) {
  if (!s) return fail("tse_comm_init: null handle");
  if (s->comm) return fail("tse_comm_init: already initialised");
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof id);
  NC(ncclCommInitRank(&s->comm, nranks, id, rank));
  s->nranks = nranks;
  s->rank = rank;
  for (const auto& c : s->cycles)
    if (c.peer < 0 || c.peer >= nranks || c.peer == rank) return fail("tse_comm_init: exchange cycle with rank %d", c.peer);
  return 0;
}

// ---- host <-> device copies -------------------------------------------------------------------
static int qdp_copy(tse_state* s, double* host, long long elem_stride, int tl, int to_device) {
  if (check_tl(tl)) return 1;
  if (!to_device && resolve_slot(s, tl)) return 1;
  const size_t per_elem = (size_t)16 * NLEV * s->Q;
  const size_t ne_chunk = s->stage_doubles / per_elem;
  const size_t tl_off = (size_t)(tl - 1) * 16 * NLEV * s->cfg.qsize_d;
  double* dev = s->qbuf[s->slot_buf[tl]];
  for (size_t e0 = 0; e0 < (size_t)s->nelem; e0 += ne_chunk) {
    const size_t n = std::min(ne_chunk, (size_t)s->nelem - e0);
    double* h = host + e0 * (size_t)elem_stride + tl_off;
    const unsigned blocks = (unsigned)((n * per_elem / 2 + 255) / 256);
    if (to_device) {
      CU(cudaMemcpy2DAsync(s->stage, per_elem * 8, h, (size_t)elem_stride * 8, per_elem * 8, n, cudaMemcpyHostToDevice, s->stream));
      k_qdp_relayout<<<blocks, 256, 0, s->stream>>>(dev, s->stage, s->d_h2i, (int)e0, (int)n, s->Q, 1);
      ++s->launches;
    } else {
      k_qdp_relayout<<<blocks, 256, 0, s->stream>>>(dev, s->stage, s->d_h2i, (int)e0, (int)n, s->Q, 0);
      ++s->launches;
      CU(cudaMemcpy2DAsync(h, (size_t)elem_stride * 8, s->stage, per_elem * 8, per_elem * 8, n, cudaMemcpyDeviceToHost, s->stream));
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));
  }
  if (to_device) s->slot_pending[tl] = 0;
  return 0;
}

int tse_copy_qdp_h2d(tse_handle s, const double* qdp, long long elem_stride, int tl) {
  if (!s || !qdp) return fail("tse_copy_qdp_h2d: null argument");
  return qdp_copy(s, const_cast<double*>(qdp), elem_stride, tl, 1);
}
int tse_copy_qdp_d2h(tse_handle s, double* qdp, long long elem_stride, int tl) {
  if (!s || !qdp) return fail("tse_copy_qdp_d2h: null argument");
  if (qdp_copy(s, qdp, elem_stride, tl, 0)) return 1;
  return check_device_error(s);  // the field may be half remapped (prim_advection_mod.F90:1323)
}

static int level_copy(tse_state* s, double* dev, double* host, long long stride, int ncomp, int host_nlev, int to_device) {
  if (!host) return 0;
  const size_t per_elem = (size_t)16 * ncomp * host_nlev;
  const size_t ne_chunk = s->stage_doubles / per_elem;
  for (size_t e0 = 0; e0 < (size_t)s->nelem; e0 += ne_chunk) {
    const size_t n = std::min(ne_chunk, (size_t)s->nelem - e0);
    double* h = host + e0 * (size_t)stride;
    const unsigned blocks = (unsigned)((n * 16 * ncomp * NLEV / 2 + 255) / 256);
    if (to_device) {
      CU(cudaMemcpy2DAsync(s->stage, per_elem * 8, h, (size_t)stride * 8, per_elem * 8, n, cudaMemcpyHostToDevice, s->stream));
      k_level_relayout<<<blocks, 256, 0, s->stream>>>(dev, s->stage, s->d_h2i + e0, (int)n, ncomp, host_nlev, 1);
      ++s->launches;
    } else {
      // host levels beyond NLEV (eta_dot_dpdn(nlev+1)) stay untouched: only the first NLEV*ncomp planes are copied back
      k_level_relayout<<<blocks, 256, 0, s->stream>>>(dev, s->stage, s->d_h2i + e0, (int)n, ncomp, host_nlev, 0);
      ++s->launches;
      CU(cudaMemcpy2DAsync(h, (size_t)stride * 8, s->stage, per_elem * 8, (size_t)16 * ncomp * NLEV * 8, n, cudaMemcpyDeviceToHost, s->stream));
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));
  }
  return 0;
}

// upload of one level field on the copy stream, chunk by chunk through the two halves of stage_up; no host synchronisation
static int level_upload_async(tse_state* s, double* dev, const double* host, long long stride, int ncomp, int* half) {
  const size_t per_elem = (size_t)16 * ncomp * NLEV;
  const size_t half_doubles = s->stage_up_doubles / 2;
  const size_t ne_chunk = half_doubles / per_elem;
  for (size_t e0 = 0; e0 < (size_t)s->nelem; e0 += ne_chunk) {
    const size_t n = std::min(ne_chunk, (size_t)s->nelem - e0);
    double* st = s->stage_up + (size_t)(*half) * half_doubles;
    *half ^= 1;
    if ((size_t)stride == per_elem)  // elements back to back on the host: one linear copy
      CU(cudaMemcpyAsync(st, host + e0 * per_elem, n * per_elem * 8, cudaMemcpyHostToDevice, s->copy_stream));
    else
      CU(cudaMemcpy2DAsync(st, per_elem * 8, host + e0 * (size_t)stride, (size_t)stride * 8, per_elem * 8, n, cudaMemcpyHostToDevice, s->copy_stream));
    const unsigned blocks = (unsigned)((n * 16 * ncomp * NLEV / 2 + 255) / 256);
    k_level_relayout<<<blocks, 256, 0, s->copy_stream>>>(dev, st, s->d_h2i + e0, (int)n, ncomp, NLEV, 1);
    ++s->launches;
    CU(cudaGetLastError());
  }
  return 0;
}

int tse_set_derived(tse_handle s, const double* vn0, long long s_vn0, const double* dp, long long s_dp, const double* eta, long long s_eta,
                    const double* omega, long long s_omega) {
  if (!s) return fail("tse_set_derived: null handle");
  if (!s) return fail("tse_set_derived: null handle");
  if (vn0 || dp) {
    // The new winds go into the second copy of vn0/dp on the copy stream while the kernels queued so far (the previous tracer
    // step) still read the first; the compute stream picks them up through an event and the two copies swap roles.  The second
    // copy was last read by the kernels queued before the previous call: ev_alt_free marks their end.
    CU(cudaStreamWaitEvent(s->copy_stream, s->ev_alt_free, 0));
    int half = 0;
    if (vn0 && level_upload_async(s, s->vn0_alt, vn0, s_vn0, 2, &half)) return 1;
    if (dp && level_upload_async(s, s->dp_alt, dp, s_dp, 1, &half)) return 1;
    CU(cudaEventRecord(s->ev_derived, s->copy_stream));
    CU(cudaStreamWaitEvent(s->stream, s->ev_derived, 0));
    if (vn0) std::swap(s->vn0, s->vn0_alt);
    if (dp) std::swap(s->dp, s->dp_alt);
    CU(cudaEventRecord(s->ev_alt_free, s->stream));
  }
  if (level_copy(s, s->eta_dot, const_cast<double*>(eta), s_eta, 1, NLEV + 1, 1)) return 1;
  if (level_copy(s, s->omega_p, const_cast<double*>(omega), s_omega, 1, NLEV, 1)) return 1;
  return 0;
}

int tse_get_derived(tse_handle s, double* divdp, long long s_divdp, double* proj, long long s_proj, double* eta, long long s_eta,
                    double* omega, long long s_omega) {
  if (!s) return fail("tse_get_derived: null handle");
  if (!s) return fail("tse_get_derived: null handle");
  if (level_copy(s, s->divdp, divdp, s_divdp, 1, NLEV, 0)) return 1;
  if (level_copy(s, s->divdp_proj, proj, s_proj, 1, NLEV, 0)) return 1;
  if (level_copy(s, s->eta_dot, eta, s_eta, 1, NLEV + 1, 0)) return 1;
  if (level_copy(s, s->omega_p, omega, s_omega, 1, NLEV, 0)) return 1;
  return 0;
}

int tse_get_wind(tse_handle s, double* vn0, long long s_vn0, double* dp, long long s_dp) {
  if (!s) return fail("tse_get_wind: null handle");
  if (level_copy(s, s->vn0, vn0, s_vn0, 2, NLEV, 0)) return 1;
  return level_copy(s, s->dp, dp, s_dp, 1, NLEV, 0);
}

int tse_get_dp3d_ps(tse_handle s, double* dp3d, long long s_dp3d, double* ps_v, long long s_ps) {
  if (!s) return fail("tse_get_dp3d_ps: null handle");
  if (level_copy(s, s->dp3d, dp3d, s_dp3d, 1, NLEV, 0)) return 1;
  if (ps_v) {
    std::vector<double> tmp((size_t)s->npad * 16);
    CU(cudaMemcpyAsync(tmp.data(), s->ps_v, tmp.size() * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    for (int eh = 0; eh < s->nelem; ++eh) std::memcpy(ps_v + (size_t)eh * s_ps, &tmp[(size_t)s->h2i[eh] * 16], 16 * 8);
  }
  CU(cudaStreamSynchronize(s->stream));
  return check_device_error(s);
}

int tse_get_qminmax(tse_handle s, double* qmin, double* qmax) {
  if (!s) return fail("tse_get_qminmax: null handle");
  const size_t per_elem = (size_t)s->Q * NLEV;
  const size_t ne_chunk = std::max<size_t>(1, s->stage_doubles / per_elem);
  for (int w = 0; w < 2; ++w) {
    double* h = w ? qmax : qmin;
    if (!h) continue;
    for (size_t e0 = 0; e0 < (size_t)s->nelem; e0 += ne_chunk) {
      const size_t n = std::min(ne_chunk, (size_t)s->nelem - e0), cnt = n * per_elem;
      k_scalar_to_host<<<(unsigned)((cnt + 255) / 256), 256, 0, s->stream>>>(w ? s->qmax : s->qmin, s->stage, s->d_h2i + e0, (int)n, s->Q);
      ++s->launches;
      CU(cudaGetLastError());
      CU(cudaMemcpyAsync(h + e0 * per_elem, s->stage, cnt * 8, cudaMemcpyDeviceToHost, s->stream));
      CU(cudaStreamSynchronize(s->stream));
    }
  }
  return 0;
}

// ---- the path ---------------------------------------------------------------------------------
int tse_precompute_divdp(tse_handle s) {
  if (!s) return fail("tse_precompute_divdp: null handle");
  if (poll_device_error(s)) return 1;
  k_divdp<<<s->ngroups * NKC / 2, DV_THREADS, DV_SMEM, s->stream>>>(s->geo, s->dvv, s->vn0, s->divdp, s->divdp_proj);
  ++s->launches;
  CU(cudaGetLastError());
  return 0;
}

// One stage of the euler_step family.  tse_euler_step and tse_advance_hypervis_scalar differ only in these numbers.
struct StageParams {
  double dt_adv;       // dt of the advective part (0: none)
  double rhs_mult_dt;  // dp = derived%dp - rhs_mult_dt * derived%divdp_proj
  double visc_coef;    // Qtens += visc_coef * dp0(k) * biharmonic / spheremp        (stage kind 2 only)
  int limiter8;        // limiter_optim_iter_full on the result
  int limiter_zero;    // limiter2d_zero on the result
  int need_bounds;     // compute qmin/qmax (neighbor_minmax)
  int store_bounds;    // write the relaxed bounds back
};
static int euler_stage(tse_state* s, int np1_qdp, int n0_qdp, int DSSopt, int rhs_multiplier, const StageParams& sp) {
  const int in = s->slot_buf[n0_qdp], in_pending = s->slot_pending[n0_qdp];
  const int other = (np1_qdp == n0_qdp) ? s->slot_buf[3 - np1_qdp] : -1;  // the untouched time level
  TileArgs a = tile_args(s);
  a.rhs_mult_dt = sp.rhs_mult_dt;
  a.dt = sp.dt_adv;
  a.visc_coef = sp.visc_coef;
  a.limiter8 = sp.limiter8;
  a.store_bounds = sp.store_bounds;
  int tmp = -1;
  if (wait_halo(s)) return 1;
  if (rhs_multiplier == 0 && sp.need_bounds) {
    // qmin/qmax = element extrema of Q = Qdp/dp, then min/max over the 8 neighbours (:764-778)
    set_src(s, a, 0, in, in_pending);
    if (launch_tile_overlapped<OP_MINMAX>(s, a, [&]() { return pack_minmax(s) || exchange(s, {xfer_minmax(s)}); })) return 1;
    if (neighbor_minmax(s)) return 1;
  } else if (rhs_multiplier == 2) {
    // biharmonic_wk_scalar_minmax (viscosity_mod.F90:353-442): lap(Q); one message per neighbour carries lap(Q) and the extrema
    // (3*nlev*qsize layers in the reference); the second laplacian runs inside the stage kernel
    tmp = pick_buffer({in, other});
    if (tmp < 0) return fail("tse_euler_step: no free tracer buffer");
    set_src(s, a, 0, in, in_pending);
    a.out = s->qbuf[tmp];
    if (launch_tile_overlapped<OP_BIHARM_PRE>(
            s, a, [&]() { return pack_tracer(s, tmp) || pack_minmax(s) || exchange(s, {xfer_tracer(s, tmp), xfer_minmax(s)}); }))
      return 1;
    // the stage kernel gathers lap(Q) of off-GPU neighbours from the ghosts: the exchange must have landed either way
    if (sp.need_bounds ? neighbor_minmax(s) : wait_halo(s)) return 1;
  }
  const int outb = pick_buffer({in, other, tmp});
  if (outb < 0) return fail("tse_euler_step: no free tracer buffer");
  a.out = s->qbuf[outb];
  double** lf = level_field(s, DSSopt);
  if (DSSopt != TSE_DSS_NO_VAR && !lf) return fail("tse_euler_step: DSSopt=%d", DSSopt);
  // the bndry_exchangeV of the stage (:923-927): Qdp(np1) and the extra level field travel in one message per neighbour while the
  // interior groups are still being computed; the unpack of Qdp happens in the next reader
  auto stage_comm = [&]() -> int {
    if (pack_tracer(s, outb)) return 1;
    if (lf) return pack_level(s, *lf) || exchange(s, {xfer_tracer(s, outb), xfer_level(s)});
    return exchange(s, {xfer_tracer(s, outb)});
  };
  {
    ScopedTimer tk(s, "k_euler_stage");
    if (rhs_multiplier == 2) {
      set_src(s, a, 0, tmp, 1);
      set_src(s, a, 1, in, in_pending);
      if (sp.limiter_zero) {
        if (launch_tile_overlapped<OP_HYPERVIS>(s, a, stage_comm)) return 1;
      } else if (launch_tile_overlapped<OP_STAGE3>(s, a, stage_comm)) return 1;
    } else {
      set_src(s, a, 0, in, in_pending);
      if (rhs_multiplier == 0) {
        if (launch_tile_overlapped<OP_STAGE1>(s, a, stage_comm)) return 1;
      } else if (launch_tile_overlapped<OP_STAGE2>(s, a, stage_comm)) return 1;
    }
    ++s->stage_launches;
  }
  s->slot_buf[np1_qdp] = outb;
  s->slot_pending[np1_qdp] = 1;
  return dss_level_field(s, DSSopt);
}

int tse_euler_step(tse_handle s, int np1_qdp, int n0_qdp, double dt, int DSSopt, int rhs_multiplier) {
  if (!s) return fail("tse_euler_step: null handle");
  if (poll_device_error(s)) return 1;
  if (check_tl(np1_qdp) || check_tl(n0_qdp)) return 1;
  if (rhs_multiplier < 0 || rhs_multiplier > 2) return fail("tse_euler_step: rhs_multiplier=%d", rhs_multiplier);
  ScopedTimer tm(s, "euler_step");
  StageParams sp{};
  sp.dt_adv = dt;
  sp.rhs_mult_dt = rhs_multiplier * dt;
  sp.visc_coef = -3.0 * dt * s->cfg.nu_q;  // rhs_viss = 3 (prim_advection_mod.F90:797,823)
  sp.limiter8 = s->cfg.limiter_option == 8 ? 1 : 0;
  sp.limiter_zero = 0;
  // Without limiter 8 nothing reads qmin/qmax (the reference computes them all the same): the fused driver skips the passes, the
  // stage-by-stage entry keeps them so that tse_get_qminmax returns what the reference holds.
  sp.need_bounds = (sp.limiter8 || !s->fused_step) ? 1 : 0;
  // Stage 2 reads the bounds stage 1 relaxed; stage 3 starts from fresh extrema (:797-806), so what stages 2 and 3 would write
  // back is never read on the path.  The fused driver skips those stores; the stage-by-stage entry keeps them for tse_get_qminmax.
  sp.store_bounds = (rhs_multiplier == 0 || !s->fused_step) ? 1 : 0;
  return euler_stage(s, np1_qdp, n0_qdp, DSSopt, rhs_multiplier, sp);
}

// advance_hypervis_scalar_cuda (cuda_mod.F90:624-718; kernels :1292-1360, :863-913, :917-928).  No executable of the reference
// calls it; it is provided as a separate entry for hosts that want HOMME's forward-in-time tracer hyperviscosity with the
// zero limiter.  Per subcycle (dt = dt2/hypervis_subcycle_q):
//   qtens = lap( dp0*Qdp/(derived%dp - dt2*divdp_proj) );  DSS;  Qdp = spheremp*Qdp - dt*nu_q*lap(rspheremp*qtens);
//   limiter2d_zero;  DSS;  Qdp *= rspheremp
// which is the stage-3 machinery (OP_BIHARM_PRE + OP_STAGE3) with the advective part switched off (dt_adv = 0), the viscous
// coefficient -dt*nu_q (dp0 applied after the second laplacian: it is constant on a level) and the zero limiter on the result;
// the last DSS is applied by the next reader, like after every stage.  nu_p = 0 branch only.
int tse_advance_hypervis_scalar(tse_handle s, int nt_qdp, double dt2) {
  if (!s) return fail("tse_advance_hypervis_scalar: null handle");
  if (poll_device_error(s)) return 1;
  if (check_tl(nt_qdp)) return 1;
  if (s->cfg.nu_q == 0.0 || s->cfg.hypervis_order != 2) return 0;  // cuda_mod.F90:656-657
  ScopedTimer tm(s, "advance_hypervis_scalar");
  const int nsub = s->cfg.hypervis_subcycle_q > 0 ? s->cfg.hypervis_subcycle_q : 1;
  StageParams sp{};
  sp.dt_adv = 0.0;
  sp.rhs_mult_dt = dt2;
  sp.visc_coef = -(dt2 / nsub) * s->cfg.nu_q;
  sp.limiter8 = 0;
  sp.limiter_zero = 1;
  sp.need_bounds = 0;
  sp.store_bounds = 0;
  for (int ic = 0; ic < nsub; ++ic)
    if (euler_stage(s, nt_qdp, nt_qdp, TSE_DSS_NO_VAR, 2, sp)) return 1;
  return 0;
}

int tse_qdp_time_avg(tse_handle s, int rkstage, int n0_qdp, int np1_qdp) {
  if (!s) return fail("tse_qdp_time_avg: null handle");
  if (poll_device_error(s)) return 1;
  if (check_tl(np1_qdp) || check_tl(n0_qdp) || n0_qdp == np1_qdp) return fail("tse_qdp_time_avg: bad time levels %d %d", n0_qdp, np1_qdp);
  if (resolve_slot(s, n0_qdp) || wait_halo(s)) return 1;
  const int in = s->slot_buf[np1_qdp], q0 = s->slot_buf[n0_qdp];
  const int outb = pick_buffer({in, q0});
  TileArgs a = tile_args(s);
  set_src(s, a, 0, q0, 0);
  set_src(s, a, 1, in, s->slot_pending[np1_qdp]);
  a.out = s->qbuf[outb];
  a.rkstage = (double)rkstage;
  launch_tile<OP_TIME_AVG>(s, a);
  CU(cudaGetLastError());
  s->slot_buf[np1_qdp] = outb;
  s->slot_pending[np1_qdp] = 0;
  return 0;
}

int tse_vertical_remap(tse_handle s, double dt, int np1, int np1_qdp) {
  if (!s) return fail("tse_vertical_remap: null handle");
  if (poll_device_error(s)) return 1;
  (void)np1;  // the device keeps a single copy of dp3d/ps_v: the one of time level np1
  if (check_tl(np1_qdp)) return 1;
  if (resolve_slot(s, np1_qdp)) return 1;
  ScopedTimer tm(s, "vertical_remap");
  RemapArgs a;
  a.q = s->qbuf[s->slot_buf[np1_qdp]];
  a.dp = s->dp; a.divdp_proj = s->divdp_proj; a.dp3d = s->dp3d; a.ps_v = s->ps_v;
  std::memcpy(a.dA, s->h_dA.data(), sizeof a.dA);
  std::memcpy(a.dB, s->h_dB.data(), sizeof a.dB);
  a.hyai0_ps0 = s->hyai0_ps0; a.dt = dt; a.Q = s->Q; a.nelem = s->nelem; a.error_flag = s->d_err;
  const int rm_threads = std::min(RM_MAX_THREADS, 32 * ((16 * s->Q + 31) / 32));
  k_vertical_remap<<<s->nelem, rm_threads, RM_SMEM, s->stream>>>(a);
  ++s->launches;
  CU(cudaGetLastError());
  return queue_error_readback(s);
}

int tse_advec_tracers_remap_rk2(tse_handle s, double dt, int nstep) {
  if (!s) return fail("tse_advec_tracers_remap_rk2: null handle");
  // TimeLevel_Qdp (time_mod.F90:85-109)
  const int qsplit = s->cfg.qsplit > 0 ? s->cfg.qsplit : 1;
  const int n0 = ((nstep / qsplit) % 2 == 0) ? 1 : 2, np1 = 3 - n0;
  ScopedTimer tm(s, "prim_advec_tracers_remap_rk2");
  if (tse_precompute_divdp(s)) return 1;
  s->fused_step = true;
  const int rc = tse_euler_step(s, np1, n0, dt / 2, TSE_DSS_DIV_VDP_AVE, 0) || tse_euler_step(s, np1, np1, dt / 2, TSE_DSS_ETA, 1) ||
                 tse_euler_step(s, np1, np1, dt / 2, TSE_DSS_OMEGA, 2);
  s->fused_step = false;
  if (rc) return 1;
  return tse_qdp_time_avg(s, 3, n0, np1);
}

int tse_dcmip_init(tse_handle s, int test_case) {
  if (!s) return fail("tse_dcmip_init: null handle");
  if (test_case != 11 && test_case != 12) return fail("tse_dcmip_init: test_case must be 11 (DCMIP 1-1) or 12 (DCMIP 1-2)");
  if (!s->have_latlon || s->hv_hyam.empty()) return fail("tse_dcmip_init: needs spherep lat/lon and hyam/hybm at tse_init");
  s->test_case = test_case;
  DcmipHostTables t;
  dcmip_fill_tables(test_case, s->hv_hyai.data(), s->hv_hybi.data(), s->hv_hyam.data(), s->hv_hybm.data(), s->ps0, t);
  auto up = [&](const double* h, const double** d) -> int {
    std::vector<double> v(h, h + NLEV);
    double* p = nullptr;
    if (upload(s, &p, v)) return 1;
    *d = p;
    return 0;
  };
  if (up(t.zm, &s->dcmip.zm) || up(t.dp_ref, &s->dcmip.dp_ref) || up(t.vm, &s->dcmip.vm) || up(t.vi, &s->dcmip.vi) ||
      up(t.dp_ic, &s->dcmip.dp_ic))
    return 1;
  {  // ps_v(t=0) = p_i(nlevp) everywhere (dcmip_wrapper_mod.F90:183): what tse_diag_qminmax divides by before the first remap
    std::vector<double> ps((size_t)s->npad * 16, t.pint[NLEV]);
    CU(cudaMemcpyAsync(s->ps_v, ps.data(), ps.size() * 8, cudaMemcpyHostToDevice, s->stream));
    CU(cudaStreamSynchronize(s->stream));
  }
  s->slot_buf[1] = 0; s->slot_buf[2] = 1;
  s->slot_pending[1] = s->slot_pending[2] = 0;
  k_dcmip_ic<<<s->ngroups * NKC, GE * 16, 0, s->stream>>>(test_case, s->nelem, s->Q, s->d_lon, s->d_lat, s->dcmip, s->qbuf[0], s->qbuf[1]);
  ++s->launches;
  CU(cudaGetLastError());
  return 0;
}

// prim_run_subcycle (prim_driver_mod.F90:701-854) with prim_step (:858-943) and prim_advance_exp (prim_advance_mod.F90:70-152)
int tse_prim_run_subcycle(tse_handle s, double tstep, int* nstep_io) {
  if (!s) return fail("tse_prim_run_subcycle: null handle");
  if (poll_device_error(s)) return 1;
  if (!s->test_case) return fail("tse_prim_run_subcycle: call tse_dcmip_init first");
  int nstep = *nstep_io;
  const int rsplit = s->cfg.rsplit > 0 ? s->cfg.rsplit : 1, qsplit = s->cfg.qsplit > 0 ? s->cfg.qsplit : 1;
  if (qsplit != 1) return fail("tse_prim_run_subcycle: qsplit=%d not supported by the device driver", qsplit);
  ScopedTimer tm(s, "prim_run");
  for (int r = 1; r <= rsplit; ++r) {
    if (r > 1) ++nstep;  // TimeLevel_update
    {
      ScopedTimer t2(s, "prim_advance_exp");
      // v(n0) was evaluated by the previous step (time 0 for the first one): winds are lagged one step
      const double t_prev = (nstep > 0 ? nstep - 1 : 0) * tstep, t_now = nstep * tstep;
      k_dcmip_wind<<<s->ngroups * NKC, GE * 16, 0, s->stream>>>(s->test_case, t_prev, t_now, s->nelem, s->d_lon, s->d_lat, s->dcmip, s->vn0,
                                                            s->dp, s->eta_dot, s->omega_p);
      ++s->launches;
      CU(cudaGetLastError());
    }
    if (tse_advec_tracers_remap_rk2(s, tstep * qsplit, nstep)) return 1;
  }
  const int np1_qdp = ((nstep / qsplit) % 2 == 0) ? 2 : 1;
  if (tse_vertical_remap(s, tstep * qsplit * rsplit, 0, np1_qdp)) return 1;
  ++nstep;
  *nstep_io = nstep;
  return 0;
}

// global tracer mass sum_e sum_k sum_ij spheremp*Qdp with an order-independent fixed-point sum
// (the repro_sum idea, repro_sum_mod.F90:216-628: bitwise identical for any element order / GPU count)
int tse_diag_mass(tse_handle s, int tl, double* mass) {
  if (!s) return fail("tse_diag_mass: null handle");
  if (check_tl(tl) || wait_halo(s)) return 1;
  if (!mass) return fail("tse_diag_mass: null argument");
  const int Q = s->Q;
  // One pass through the tile pipeline (OP_MASS): every plane's J = sum spheremp*Qdp is added as a two-limb fixed-point number
  // scaled by 2^shift[q], and the largest |J| is tracked alongside.  The shift of the previous call is reused as long as the
  // largest |J| stays inside its window (2^26 <= |J|max * 2^shift < 2^38: no overflow with up to 2^23 planes, >= 66 significant
  // bits); otherwise it is re-centred and the pass repeated (first call, or a tracer whose mass changed by orders of magnitude).
  // Max and sums are integer all-reduces, so every rank takes the same decision and the result is bitwise independent of the
  // element order and of the number of GPUs.
  TileArgs a = tile_args(s);
  set_src(s, a, 0, s->slot_buf[tl], s->slot_pending[tl]);
  a.mass_acc = s->d_acc;
  a.mass_maxbits = s->d_maxbits;
  a.mass_shift = s->d_shift;
  std::vector<unsigned long long> mb(Q), mb_all((size_t)MASS_REP * Q);
  std::vector<long long> acc(2 * Q), acc_all((size_t)2 * MASS_REP * Q);
  for (int attempt = 0; attempt < 2; ++attempt) {
    CU(cudaMemsetAsync(s->d_maxbits, 0, sizeof(unsigned long long) * MASS_REP * Q, s->stream));
    CU(cudaMemsetAsync(s->d_acc, 0, sizeof(long long) * 2 * MASS_REP * Q, s->stream));
    launch_tile<OP_MASS>(s, a);
    CU(cudaGetLastError());
    if (s->comm) {
      // doubles >= 0 order like their bit patterns: the global max is an integer max
      NC(ncclAllReduce(s->d_maxbits, s->d_maxbits, (size_t)MASS_REP * Q, ncclUint64, ncclMax, s->comm, s->stream));
      NC(ncclAllReduce(s->d_acc, s->d_acc, (size_t)2 * MASS_REP * Q, ncclInt64, ncclSum, s->comm, s->stream));
    }
    CU(cudaMemcpyAsync(mb_all.data(), s->d_maxbits, sizeof(unsigned long long) * MASS_REP * Q, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemcpyAsync(acc_all.data(), s->d_acc, sizeof(long long) * 2 * MASS_REP * Q, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    std::fill(mb.begin(), mb.end(), 0ull);
    std::fill(acc.begin(), acc.end(), 0ll);
    for (int r = 0; r < MASS_REP; ++r)
      for (int q = 0; q < Q; ++q) {
        mb[q] = std::max(mb[q], mb_all[(size_t)r * Q + q]);
        // wrapping adds, like the device atomics: the true sums fit (the window below), so the wrapped total is exact
        acc[2 * q] = (long long)((unsigned long long)acc[2 * q] + (unsigned long long)acc_all[2 * ((size_t)r * Q + q)]);
        acc[2 * q + 1] = (long long)((unsigned long long)acc[2 * q + 1] + (unsigned long long)acc_all[2 * ((size_t)r * Q + q) + 1]);
      }
    bool fits = true;
    for (int q = 0; q < Q; ++q) {
      double mx;
      std::memcpy(&mx, &mb[q], 8);
      const int ideal = 37 - (mx > 0 ? std::ilogb(mx) : 0);  // |J|max * 2^ideal in [2^37, 2^38)
      if (mx > 0 && (s->mass_shift[q] > ideal || s->mass_shift[q] < ideal - 11)) fits = false;
      if (!fits) break;
    }
    if (fits) break;
    if (attempt == 1) return fail("tse_diag_mass: fixed-point window did not settle");
    for (int q = 0; q < Q; ++q) {
      double mx;
      std::memcpy(&mx, &mb[q], 8);
      s->mass_shift[q] = 37 - (mx > 0 ? std::ilogb(mx) : 0);
    }
    CU(cudaMemcpyAsync(s->d_shift, s->mass_shift.data(), sizeof(int) * Q, cudaMemcpyHostToDevice, s->stream));
  }
  for (int q = 0; q < Q; ++q) {
    const long double tot = (long double)acc[2 * q] + std::ldexp((long double)acc[2 * q + 1], -40);
    mass[q] = (double)std::ldexp(tot, -s->mass_shift[q]);
  }
  return check_device_error(s);
}
int tse_diag_field_hash(tse_handle s, int tl, unsigned long long* hash) {
  if (!s || !hash) return fail("tse_diag_field_hash: null argument");
  if (check_tl(tl) || wait_halo(s)) return 1;
  const DssView v = view(s, s->slot_buf[tl], s->slot_pending[tl]);
  const int Q = s->Q;
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(s->d_acc);
  CU(cudaMemsetAsync(acc, 0, sizeof(unsigned long long) * Q, s->stream));
  k_field_hash<<<plane_grid(s), GPL * QPB, 0, s->stream>>>(s->geo, v, s->d_gkey, acc);
  ++s->launches;
  CU(cudaGetLastError());
  if (s->comm) NC(ncclAllReduce(acc, acc, Q, ncclUint64, ncclSum, s->comm, s->stream));  // wraps modulo 2^64 like the local sums
  CU(cudaMemcpyAsync(hash, acc, sizeof(unsigned long long) * Q, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return check_device_error(s);
}
int tse_diag_qminmax(tse_handle s, int tl, double* qmin, double* qmax) {
  if (!s) return fail("tse_diag_qminmax: null handle");
  if (check_tl(tl) || wait_halo(s)) return 1;
  const DssView v = view(s, s->slot_buf[tl], s->slot_pending[tl]);
  const int Q = s->Q;
  // d_acc doubles as the pair of ordered-integer accumulators: [0,Q) min, [Q,2Q) max
  unsigned long long* acc = reinterpret_cast<unsigned long long*>(s->d_acc);
  CU(cudaMemsetAsync(acc, 0xff, sizeof(unsigned long long) * Q, s->stream));
  CU(cudaMemsetAsync(acc + Q, 0, sizeof(unsigned long long) * Q, s->stream));
  k_q_minmax<<<plane_grid(s), GPL * QPB, 0, s->stream>>>(s->geo, v, s->d_dA, s->d_dB, s->ps_v, acc, acc + Q);
  ++s->launches;
  CU(cudaGetLastError());
  if (s->comm) {
    NC(ncclAllReduce(acc, acc, Q, ncclUint64, ncclMin, s->comm, s->stream));
    NC(ncclAllReduce(acc + Q, acc + Q, Q, ncclUint64, ncclMax, s->comm, s->stream));
  }
  std::vector<unsigned long long> h(2 * Q);
  CU(cudaMemcpyAsync(h.data(), acc, sizeof(unsigned long long) * 2 * Q, cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  auto back = [](unsigned long long b) {
    b = (b >> 63) ? (b & 0x7fffffffffffffffull) : ~b;
    double x;
    std::memcpy(&x, &b, 8);
    return x;
  };
  for (int q = 0; q < Q; ++q) {
    qmin[q] = back(h[q]);
    qmax[q] = back(h[Q + q]);
  }
  return check_device_error(s);
}

#ifdef TSE_EXP_LIMSTATS
extern "C" int tse_exp_limiter_stats(unsigned long long* out4, int reset) {
  cudaDeviceSynchronize();
  if (out4 && cudaMemcpyFromSymbol(out4, tse::g_lim_stats, 32) != cudaSuccess) return 1;
  if (reset) {
    const unsigned long long z[4] = {0, 0, 0, 0};
    if (cudaMemcpyToSymbol(tse::g_lim_stats, z, 32) != cudaSuccess) return 1;
  }
  return 0;
}
#endif
int tse_debug_limiter(int n, double* ptens_w, const double* sphweights, const double* dpmass, double* minp, double* maxp) {
  if (n <= 0) return 0;
  if (!ptens_w || !sphweights || !dpmass || !minp || !maxp) return fail("tse_debug_limiter: null argument");
  double *d_y = nullptr, *d_s = nullptr, *d_d = nullptr, *d_mn = nullptr, *d_mx = nullptr;
  const size_t nb = (size_t)n * 16 * 8;
  int rc = 0;
  auto ok = [&](cudaError_t e, const char* what) {
    if (e != cudaSuccess && !rc) rc = fail("tse_debug_limiter: %s: %s", what, cudaGetErrorString(e));
    return e == cudaSuccess;
  };
  if (ok(cudaMalloc(&d_y, nb), "cudaMalloc") && ok(cudaMalloc(&d_s, nb), "cudaMalloc") && ok(cudaMalloc(&d_d, nb), "cudaMalloc") &&
      ok(cudaMalloc(&d_mn, (size_t)n * 8), "cudaMalloc") && ok(cudaMalloc(&d_mx, (size_t)n * 8), "cudaMalloc") &&
      ok(cudaMemcpy(d_y, ptens_w, nb, cudaMemcpyHostToDevice), "h2d") && ok(cudaMemcpy(d_s, sphweights, nb, cudaMemcpyHostToDevice), "h2d") &&
      ok(cudaMemcpy(d_d, dpmass, nb, cudaMemcpyHostToDevice), "h2d") && ok(cudaMemcpy(d_mn, minp, (size_t)n * 8, cudaMemcpyHostToDevice), "h2d") &&
      ok(cudaMemcpy(d_mx, maxp, (size_t)n * 8, cudaMemcpyHostToDevice), "h2d")) {
    k_debug_limiter<<<(n + GPL - 1) / GPL, GPL>>>(n, d_y, d_s, d_d, d_mn, d_mx);
    ok(cudaGetLastError(), "launch");
    ok(cudaDeviceSynchronize(), "kernel");
    ok(cudaMemcpy(ptens_w, d_y, nb, cudaMemcpyDeviceToHost), "d2h");
    ok(cudaMemcpy(minp, d_mn, (size_t)n * 8, cudaMemcpyDeviceToHost), "d2h");
    ok(cudaMemcpy(maxp, d_mx, (size_t)n * 8, cudaMemcpyDeviceToHost), "d2h");
  }
  cudaFree(d_y); cudaFree(d_s); cudaFree(d_d); cudaFree(d_mn); cudaFree(d_mx);
  return rc;
}

double tse_timer_ms(tse_handle s, const char* name) {
  if (!s || !name) return -1.0;
  resolve_timers(s);
  auto it = s->timers.find(name);
  return it == s->timers.end() ? -1.0 : it->second;
}
int tse_timer_reset(tse_handle s) {
  if (!s) return fail("tse_timer_reset: null handle");
  resolve_timers(s);
  s->timers.clear();
  return 0;
}
long long tse_launch_count(tse_handle s) { return s ? s->launches : -1; }
long long tse_stage_launch_count(tse_handle s) { return s ? s->stage_launches : -1; }
long long tse_device_bytes(tse_handle s) { return s ? s->dev_bytes : -1; }
long long tse_halo_bytes(tse_handle s) { return s ? s->halo_bytes : -1; }

int tse_mark(tse_handle s, int slot) {
  if (!s) return fail("tse_mark: null handle");
  if (slot < 0 || slot >= 16) return fail("tse_mark: slot %d", slot);
  if (!s->marks[slot]) CU(cudaEventCreate(&s->marks[slot]));
  CU(cudaEventRecord(s->marks[slot], s->stream));
  return 0;
}
double tse_mark_elapsed_ms(tse_handle s, int a, int b) {
  if (!s) return -1.0;
  if (a < 0 || a >= 16 || b < 0 || b >= 16 || !s->marks[a] || !s->marks[b]) return -1.0;
  if (cudaEventSynchronize(s->marks[b]) != cudaSuccess) return -1.0;
  float ms = 0;
  if (cudaEventElapsedTime(&ms, s->marks[a], s->marks[b]) != cudaSuccess) return -1.0;
  return ms;
}

}  // extern "C"
