/* tse.h -- C ABI of the B200-native tracer-advection path of transport_se.
 *
 * This is the drop-in boundary: the entry points are what a Fortran host binds
 * through ISO_C_BINDING in place of the reference's `#if USE_CUDA_FORTRAN`
 * hooks (citations are relative to the reference tree):
 *
 *   tse_init                      <- cuda_mod_init           src/share/cuda_mod.F90:168,
 *                                                              called at src/share/prim_driver_mod.F90:686-689
 *   tse_copy_qdp_h2d / _d2h       <- copy_qdp_h2d / _d2h     src/share/cuda_mod.F90:429,451,
 *                                                              called at prim_driver_mod.F90:781-784,798-801
 *   tse_euler_step                <- euler_step_cuda         src/share/cuda_mod.F90:473,
 *                                                              called at src/share/prim_advection_mod.F90:715-718
 *   tse_qdp_time_avg              <- qdp_time_avg_cuda       src/share/cuda_mod.F90:601 (prim_advection_mod.F90:653-656)
 *   tse_vertical_remap            <- vertical_remap_cuda     src/share/cuda_mod.F90:1436 (prim_advection_mod.F90:1279-1282)
 *   tse_precompute_divdp          <- the divdp loop of Prim_Advec_Tracers_remap_rk2, prim_advection_mod.F90:614-623
 *   tse_advec_tracers_remap_rk2   <- Prim_Advec_Tracers_remap_rk2, prim_advection_mod.F90:579-640 (fused convenience)
 *   tse_set_derived / tse_get_*   <- the per-stage H2D of dp / vstar in euler_step_cuda (cuda_mod.F90:526-552)
 *   tse_diag_mass / _minmax       <- prim_printstate "qv=" / "Q,Q diss" lines, src/share/prim_state_mod.F90:184-207,352-385
 *   tse_comm_init                 <- the Schedule_t/Cycle_t exchange of bndry_exchangeV, src/share/bndry_mod.F90:21-126
 *
 * Conventions: every entry returns 0 on success and a nonzero status otherwise
 * (the Fortran shim turns it into abortmp(tse_last_error())).  All pointers are
 * HOST pointers; all arrays are column-major exactly as the Fortran types lay
 * them out, so a shim passes c_loc(elem(1)%state%Qdp) and the element stride
 * storage_size(elem(1))/64 (in doubles).  np = 4 and nlev = 72 are compile-time
 * (reference dimensions_mod.F90:12-28); qsize is runtime.  Time-level and
 * DSSopt arguments are 1-based / valued as in the reference.
 *
 * There is no CPU fallback: every compute entry fails if no CUDA device is usable.
 */
#ifndef TSE_H
#define TSE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tse_state* tse_handle;

/* prim_advection_mod.F90:454-457 */
#define TSE_DSS_ETA 1
#define TSE_DSS_OMEGA 2
#define TSE_DSS_DIV_VDP_AVE 3
#define TSE_DSS_NO_VAR (-1)

/* control_mod / dimensions_mod values the path depends on */
typedef struct {
  int ne;                  /* elements per cube edge (informational) */
  int nelemd;              /* elements owned by this rank (dimensions_mod nelemd) */
  int qsize;               /* advected tracers */
  int qsize_d;             /* extent of the tracer dimension of state%Qdp on the host */
  int nlev;                /* must be 72 */
  int np;                  /* must be 4 */
  int rsplit, qsplit;      /* 3, 1 in the shipped namelists */
  int limiter_option;      /* must be 8 (limiter_optim_iter_full) */
  int hypervis_order;      /* 2 */
  int hypervis_subcycle_q; /* must be 1 with limiter 8 (namelist_mod.F90:688-692) */
  int vert_remap_q_alg;    /* != 2: mirrored PPM boundary cells */
  double nu_q;             /* tracer hyperviscosity coefficient */
  int device;              /* CUDA device ordinal, -1 = current */
} tse_config;

/* element_t fields (element_mod.F90:112-221), one entry per local element, elements contiguous */
typedef struct {
  const double* spheremp;  /* [nelemd][np*np] */
  const double* rspheremp; /* [nelemd][np*np] */
  const double* metdet;    /* [nelemd][np*np] */
  const double* rmetdet;   /* [nelemd][np*np] */
  const double* Dinv;      /* [nelemd][np*np][2*2], Fortran Dinv(2,2,np,np) */
  const double* lat;       /* [nelemd][np*np] spherep%lat (may be NULL if prescribed winds are not used) */
  const double* lon;       /* [nelemd][np*np] spherep%lon */
} tse_geometry;

/* EdgeDescriptor_t (edge_mod.F90:31-43) + Schedule_t cycles (schedtype_mod.F90:7-29).
 * Direction index: west,east,south,north,swest,seast,nwest,neast (control_mod.F90:173-181), 0-based.
 * Offsets are 0-based positions in the horizontal dimension of the edge buffer, -1 = no neighbour. */
typedef struct {
  const int* putmapP;  /* [nelemd][8] */
  const int* getmapP;  /* [nelemd][8] */
  const int* reverse;  /* [nelemd][8] 0/1 */
  int nbuf;            /* horizontal size of the edge buffer */
  const int* sfc_index; /* [nelemd] GridVertex%SpaceCurve; NULL = keep the given order (only used for memory placement) */
  int ncycles;         /* neighbour ranks */
  const int* cyc_rank; /* [ncycles] */
  const int* cyc_ptr;  /* [ncycles] ptrP */
  const int* cyc_len;  /* [ncycles] lengthP */
} tse_connectivity;

/* hvcoord_t (hybvcoord_mod.F90:18-30) */
typedef struct {
  const double* hyai; /* [nlev+1] */
  const double* hybi; /* [nlev+1] */
  const double* hyam; /* [nlev]   (prescribed winds only; may be NULL) */
  const double* hybm; /* [nlev] */
  double ps0;
} tse_hvcoord;

const char* tse_last_error(void);
int tse_device_count(void);

int tse_init(const tse_config* cfg, const tse_geometry* geom, const tse_connectivity* conn, const tse_hvcoord* hv,
             const double* dvv /* deriv%Dvv(np,np) */, tse_handle* out);
int tse_finalize(tse_handle h);
int tse_synchronize(tse_handle h);

/* multi-GPU: one rank per GPU, elements split by the host along the space-filling curve (the connectivity of tse_init carries
 * the exchange cycles).  id128 is an ncclUniqueId created on rank 0 (tse_comm_unique_id) and broadcast by the host (MPI_Bcast in
 * a Fortran host, torch.distributed in the Python harness).  Must be called before the first compute entry when ncycles > 0. */
int tse_comm_unique_id(void* id128);
int tse_comm_init(tse_handle h, int nranks, int rank, const void* id128);

/* state%Qdp(np,np,nlev,qsize_d,2) of element 1; elem_stride in doubles between consecutive elements; tl = 1|2 */
int tse_copy_qdp_h2d(tse_handle h, const double* qdp, long long elem_stride, int tl);
int tse_copy_qdp_d2h(tse_handle h, double* qdp, long long elem_stride, int tl);

/* derived%vn0(np,np,2,nlev), derived%dp(np,np,nlev), derived%eta_dot_dpdn(np,np,nlev+1), derived%omega_p(np,np,nlev);
 * any pointer may be NULL (field left untouched).  Strides in doubles between consecutive elements.
 * vn0 and dp are uploaded asynchronously (own stream, second device copy) so that the transfer for step n+1 overlaps the kernels
 * of step n; device work queued later sees the new values.  From page-locked host memory the copy is still in flight when the
 * call returns: do not overwrite those buffers before the next blocking entry (tse_synchronize, tse_get_*, tse_diag_*, tse_copy_*). */
int tse_set_derived(tse_handle h, const double* vn0, long long s_vn0, const double* dp, long long s_dp,
                    const double* eta_dot_dpdn, long long s_eta, const double* omega_p, long long s_omega);
/* derived%divdp, divdp_proj, eta_dot_dpdn (first nlev levels), omega_p back to the host; NULL = skip */
int tse_get_derived(tse_handle h, double* divdp, long long s_divdp, double* divdp_proj, long long s_proj,
                    double* eta_dot_dpdn, long long s_eta, double* omega_p, long long s_omega);
/* state%dp3d(np,np,nlev) and state%ps_v(np,np) of time level np1 as left by tse_vertical_remap */
int tse_get_dp3d_ps(tse_handle h, double* dp3d, long long s_dp3d, double* ps_v, long long s_ps);
/* qmin/qmax(nlev,qsize) per element (prim_advection_mod.F90:461), stride nlev*qsize */
int tse_get_qminmax(tse_handle h, double* qmin, double* qmax);

int tse_precompute_divdp(tse_handle h);
int tse_euler_step(tse_handle h, int np1_qdp, int n0_qdp, double dt, int DSSopt, int rhs_multiplier);
int tse_qdp_time_avg(tse_handle h, int rkstage, int n0_qdp, int np1_qdp);
int tse_vertical_remap(tse_handle h, double dt, int np1, int np1_qdp);
int tse_advec_tracers_remap_rk2(tse_handle h, double dt, int nstep);
/* advance_hypervis_scalar_cuda(edgeAdv,elem,hvcoord,hybrid,deriv,nt,nt_qdp,nets,nete,dt2) (cuda_mod.F90:624-718): hypervis_subcycle_q
 * subcycles of Qdp(nt_qdp) += -dt*nu_q*biharmonic(dp0*Qdp/dp) with dp = derived%dp - dt2*derived%divdp_proj, each followed by
 * limiter2d_zero (:863-913) and the DSS.  No executable of the reference calls this routine (its CPU path applies the tracer
 * hyperviscosity inside the third euler_step stage); it is a separate entry, never called by the other entries.  nu_p = 0 only. */
int tse_advance_hypervis_scalar(tse_handle h, int nt_qdp, double dt2);

/* Device-side test-case driver (prim_advance_exp + prim_step + prim_run_subcycle sequencing,
 * prim_advance_mod.F90:70-152, prim_driver_mod.F90:701-943) for test_case 11 (DCMIP 1-1) / 12 (DCMIP 1-2). */
int tse_dcmip_init(tse_handle h, int test_case);                 /* prim_init2 part: IC for Qdp(1:2), winds at t=0 */
int tse_prim_run_subcycle(tse_handle h, double tstep, int* nstep /* in/out tl%nstep */);

/* diagnostics: per-tracer global mass sum(Qdp*spheremp) of time level tl (order-independent fixed-point sum,
 * repro_sum_mod.F90:216-628) and min/max of Qdp/dp */
int tse_diag_mass(tse_handle h, int tl, double* mass /* [qsize] */);
int tse_diag_qminmax(tse_handle h, int tl, double* qmin /* [qsize] */, double* qmax /* [qsize] */);
/* Per-tracer 64-bit fingerprint of Qdp(tl): wrapping sum over all (element, level, node) of a mix of the value's bit pattern and its
 * global position (GridVertex%SpaceCurve of the element, level, node).  Independent of the partition and of the element order:
 * equal fingerprints on 1, 2, 4, 8 GPUs mean the fields are bit-for-bit equal (the reference's claim, README:46-47). */
int tse_diag_field_hash(tse_handle h, int tl, unsigned long long* hash /* [qsize] */);

/* Verification hook (the reference has no counterpart): runs limiter_optim_iter_full (prim_advection_mod.F90:976-1094) exactly as the
 * stage kernels call it on n independent 4x4 planes.  ptens_w[n][16] in: ptens (tracer mass, the reference's argument); out:
 * sphweights*ptens_limited (what euler_step stores, :905).  minp/maxp[n] in/out as in the reference.  Needs no handle. */
int tse_debug_limiter(int n, double* ptens_w, const double* sphweights, const double* dpmass, double* minp, double* maxp);

/* timers: CUDA-event time (ms) accumulated under the reference's GPTL timer names
 * ("prim_advec_tracers_remap_rk2", "euler_step", "vertical_remap", "prim_advance_exp"); returns <0 if unknown */
double tse_timer_ms(tse_handle h, const char* name);
int tse_timer_reset(tse_handle h);
/* CUDA events on the handle's own stream (torch.cuda.Event only sees torch's stream): record slot 0..15, elapsed ms a->b */
int tse_mark(tse_handle h, int slot);
double tse_mark_elapsed_ms(tse_handle h, int a, int b);
/* derived%vn0 / derived%dp as currently held on the device (e.g. after the device-side prim_advance_exp) */
int tse_get_wind(tse_handle h, double* vn0, long long s_vn0, double* dp, long long s_dp);
long long tse_stage_launch_count(tse_handle h);
/* bytes this rank has sent through the halo exchange so far */
long long tse_halo_bytes(tse_handle h);
/* number of kernel launches issued by this handle so far */
long long tse_launch_count(tse_handle h);
/* device bytes allocated by this handle */
long long tse_device_bytes(tse_handle h);

#ifdef __cplusplus
}
#endif
#endif /* TSE_H */
