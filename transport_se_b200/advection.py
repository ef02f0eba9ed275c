"""Host-side mirror of the reference's tracer-advection interface over the C ABI (include/tse.h).

Method names, argument order/meaning and the 1-based time-level / DSSopt conventions follow
reference src/share/prim_advection_mod.F90 (Prim_Advec_Tracers_remap_rk2 :579-640, euler_step :667-970,
qdp_time_avg :645-662, vertical_remap :1242-1330) and the USE_CUDA_FORTRAN hooks of src/share/cuda_mod.F90.
Host arrays use the reference's element-major Fortran layouts (C order, (i,j) fastest = last axis of 16):

    Qdp[e, tl(2), q(qsize_d), k(72), 16]      vn0[e, k, c(2), 16]      dp / divdp / omega_p[e, k, 16]
    eta_dot_dpdn[e, 73, 16]                   ps_v[e, 16]              qmin / qmax[e, q, k]

All compute runs in libtse_cuda.so (hand-written sm_100a kernels).  There is no CPU fallback:
construction raises if the library or a CUDA device is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

DSSeta, DSSomega, DSSdiv_vdp_ave, DSSno_var = 1, 2, 3, -1  # prim_advection_mod.F90:454-457


class TseConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("ne", "nelemd", "qsize", "qsize_d", "nlev", "np", "rsplit", "qsplit", "limiter_option",
                                       "hypervis_order", "hypervis_subcycle_q", "vert_remap_q_alg")] + \
               [("nu_q", C.c_double), ("device", C.c_int)]


class TseGeometry(C.Structure):
    _fields_ = [(n, _dp) for n in ("spheremp", "rspheremp", "metdet", "rmetdet", "Dinv", "lat", "lon")]


class TseConnectivity(C.Structure):
    _fields_ = [("putmapP", _ip), ("getmapP", _ip), ("reverse", _ip), ("nbuf", C.c_int), ("sfc_index", _ip), ("ncycles", C.c_int),
                ("cyc_rank", _ip), ("cyc_ptr", _ip), ("cyc_len", _ip)]


class TseHvcoord(C.Structure):
    _fields_ = [("hyai", _dp), ("hybi", _dp), ("hyam", _dp), ("hybm", _dp), ("ps0", C.c_double)]


EXPORTS = ["tse_last_error", "tse_device_count", "tse_init", "tse_finalize", "tse_synchronize", "tse_comm_unique_id", "tse_comm_init",
           "tse_copy_qdp_h2d", "tse_copy_qdp_d2h", "tse_set_derived", "tse_get_derived", "tse_get_dp3d_ps", "tse_get_qminmax",
           "tse_precompute_divdp", "tse_euler_step", "tse_qdp_time_avg", "tse_vertical_remap", "tse_advec_tracers_remap_rk2",
           "tse_dcmip_init", "tse_prim_run_subcycle", "tse_diag_mass", "tse_diag_qminmax", "tse_timer_ms", "tse_launch_count",
           "tse_device_bytes", "tse_timer_reset", "tse_mark", "tse_mark_elapsed_ms", "tse_get_wind", "tse_stage_launch_count", "tse_halo_bytes",
           "tse_debug_limiter", "tse_diag_field_hash", "tse_advance_hypervis_scalar"]


def _preload_bundled_nccl():
    """libtse_cuda.so needs libnccl.so.2.  When this process also uses PyTorch (the harness does, for torch.distributed), the
    NCCL that gets loaded first wins for both; PyTorch needs the newer one it ships (nvidia/nccl/lib).  Load that one first if it
    is there, so that the order in which a test imports torch and this module does not matter.  Stand-alone C / Fortran hosts
    simply use the system library."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
            path = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(path):
                C.CDLL(path, mode=C.RTLD_GLOBAL)
                return
    except Exception:
        pass


def cuda_lib():
    """Load libtse_cuda.so (built in-tree by transport_se_b200._build.build_cuda / __graft_entry__.build)."""
    global _LIB
    if _LIB is None:
        _preload_bundled_nccl()
        path = os.environ.get("TSE_CUDA_LIB") or os.path.join(_HERE, "libtse_cuda.so")  # override: kernel-variant experiments (tools/)
        if not os.path.exists(path):
            raise RuntimeError("libtse_cuda.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback)")
        L = C.CDLL(path)
        vp, ll, d, i = C.c_void_p, C.c_longlong, C.c_double, C.c_int
        L.tse_last_error.restype = C.c_char_p
        L.tse_init.argtypes = [C.POINTER(TseConfig), C.POINTER(TseGeometry), C.POINTER(TseConnectivity), C.POINTER(TseHvcoord), _dp,
                               C.POINTER(vp)]
        L.tse_finalize.argtypes = [vp]
        L.tse_synchronize.argtypes = [vp]
        L.tse_comm_unique_id.argtypes = [vp]
        L.tse_comm_init.argtypes = [vp, i, i, vp]
        L.tse_copy_qdp_h2d.argtypes = [vp, _dp, ll, i]
        L.tse_copy_qdp_d2h.argtypes = [vp, _dp, ll, i]
        L.tse_set_derived.argtypes = [vp, _dp, ll, _dp, ll, _dp, ll, _dp, ll]
        L.tse_get_derived.argtypes = [vp, _dp, ll, _dp, ll, _dp, ll, _dp, ll]
        L.tse_get_dp3d_ps.argtypes = [vp, _dp, ll, _dp, ll]
        L.tse_get_qminmax.argtypes = [vp, _dp, _dp]
        L.tse_precompute_divdp.argtypes = [vp]
        L.tse_euler_step.argtypes = [vp, i, i, d, i, i]
        L.tse_qdp_time_avg.argtypes = [vp, i, i, i]
        L.tse_vertical_remap.argtypes = [vp, d, i, i]
        L.tse_advec_tracers_remap_rk2.argtypes = [vp, d, i]
        L.tse_advance_hypervis_scalar.argtypes = [vp, i, d]
        L.tse_dcmip_init.argtypes = [vp, i]
        L.tse_prim_run_subcycle.argtypes = [vp, d, _ip]
        L.tse_diag_mass.argtypes = [vp, i, _dp]
        L.tse_diag_qminmax.argtypes = [vp, i, _dp, _dp]
        L.tse_timer_ms.argtypes = [vp, C.c_char_p]
        L.tse_timer_ms.restype = d
        L.tse_launch_count.argtypes = [vp]
        L.tse_launch_count.restype = ll
        L.tse_device_bytes.argtypes = [vp]
        L.tse_device_bytes.restype = ll
        L.tse_stage_launch_count.argtypes = [vp]
        L.tse_stage_launch_count.restype = ll
        L.tse_timer_reset.argtypes = [vp]
        L.tse_mark.argtypes = [vp, i]
        L.tse_mark_elapsed_ms.argtypes = [vp, i, i]
        L.tse_mark_elapsed_ms.restype = d
        L.tse_get_wind.argtypes = [vp, _dp, ll, _dp, ll]
        L.tse_halo_bytes.argtypes = [vp]
        L.tse_halo_bytes.restype = ll
        L.tse_diag_field_hash.argtypes = [vp, i, C.POINTER(C.c_ulonglong)]
        L.tse_debug_limiter.argtypes = [i, _dp, _dp, _dp, _dp, _dp]
        _LIB = L
    return _LIB


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _stride(a):
    return 0 if a is None else a.strides[0] // 8


class TseError(RuntimeError):
    pass


class TracerAdvection:
    """One rank's tracer-advection state, resident in HBM (the equivalent of cuda_mod's module state)."""

    def __init__(self, mesh, view, vcoord, qsize, qsize_d=None, nu_q=0.0, rsplit=3, qsplit=1, limiter_option=8, hypervis_subcycle_q=1,
                 vert_remap_q_alg=0, device=-1, ps0=100000.0):
        L = cuda_lib()
        self._L = L
        self._h = C.c_void_p()
        g = view.gid
        c = np.ascontiguousarray
        self.nelemd, self.qsize, self.qsize_d, self.nlev = len(g), qsize, qsize_d or qsize, 72
        cfg = TseConfig(ne=mesh.ne, nelemd=self.nelemd, qsize=qsize, qsize_d=self.qsize_d, nlev=72, np=4, rsplit=rsplit, qsplit=qsplit,
                        limiter_option=limiter_option, hypervis_order=2, hypervis_subcycle_q=hypervis_subcycle_q,
                        vert_remap_q_alg=vert_remap_q_alg, nu_q=nu_q, device=device)
        keep = [c(mesh.spheremp[g]), c(mesh.rspheremp[g]), c(mesh.metdet[g]), c(mesh.rmetdet[g]), c(mesh.Dinv[g]), c(mesh.lat[g]),
                c(mesh.lon[g])]
        geom = TseGeometry(*[_p(a) for a in keep])
        ints = [c(view.putmap), c(view.getmap), c(view.reverse), c(mesh.sfc[g]), c(view.cyc_rank), c(view.cyc_ptr), c(view.cyc_len)]
        conn = TseConnectivity(_p(ints[0], C.c_int), _p(ints[1], C.c_int), _p(ints[2], C.c_int), view.nbuf, _p(ints[3], C.c_int),
                               view.ncycles, _p(ints[4], C.c_int), _p(ints[5], C.c_int), _p(ints[6], C.c_int))
        hv = [c(vcoord[k]) for k in ("hyai", "hybi", "hyam", "hybm")]
        hvc = TseHvcoord(_p(hv[0]), _p(hv[1]), _p(hv[2]), _p(hv[3]), ps0)
        dvv = c(mesh.dvv)
        self._ck(L.tse_init(C.byref(cfg), C.byref(geom), C.byref(conn), C.byref(hvc), _p(dvv), C.byref(self._h)))

    def _ck(self, rc):
        if rc != 0:
            raise TseError(self._L.tse_last_error().decode())

    def close(self):
        if self._h:
            self._L.tse_finalize(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, dist, rank, world):
        """Multi-GPU: broadcast rank 0's ncclUniqueId over torch.distributed (plumbing only) and create the halo communicator."""
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            self._ck(self._L.tse_comm_unique_id(buf))
        box = [bytes(buf)]
        dist.broadcast_object_list(box, src=0)
        idb = (C.c_ubyte * 128).from_buffer_copy(box[0])
        self._ck(self._L.tse_comm_init(self._h, world, rank, idb))

    @property
    def halo_bytes(self):
        return self._L.tse_halo_bytes(self._h)

    # --- cuda_mod hooks -------------------------------------------------------------------
    def copy_qdp_h2d(self, Qdp, tl):
        """Qdp[e, 2, qsize_d, 72, 16] (state%Qdp of every element); uploads time level tl (1|2)."""
        assert Qdp.dtype == np.float64 and Qdp.shape[1:] == (2, self.qsize_d, 72, 16) and Qdp[0].flags.c_contiguous
        self._ck(self._L.tse_copy_qdp_h2d(self._h, _p(Qdp), _stride(Qdp), tl))

    def copy_qdp_d2h(self, Qdp, tl):
        assert Qdp.dtype == np.float64 and Qdp.shape[1:] == (2, self.qsize_d, 72, 16) and Qdp[0].flags.c_contiguous
        self._ck(self._L.tse_copy_qdp_d2h(self._h, _p(Qdp), _stride(Qdp), tl))

    def set_derived(self, vn0=None, dp=None, eta_dot_dpdn=None, omega_p=None):
        self._ck(self._L.tse_set_derived(self._h, _p(vn0), _stride(vn0), _p(dp), _stride(dp), _p(eta_dot_dpdn), _stride(eta_dot_dpdn),
                                         _p(omega_p), _stride(omega_p)))

    def get_derived(self, divdp=None, divdp_proj=None, eta_dot_dpdn=None, omega_p=None):
        self._ck(self._L.tse_get_derived(self._h, _p(divdp), _stride(divdp), _p(divdp_proj), _stride(divdp_proj), _p(eta_dot_dpdn),
                                         _stride(eta_dot_dpdn), _p(omega_p), _stride(omega_p)))

    def get_dp3d_ps(self, dp3d=None, ps_v=None):
        self._ck(self._L.tse_get_dp3d_ps(self._h, _p(dp3d), _stride(dp3d), _p(ps_v), _stride(ps_v)))

    def get_qminmax(self):
        qmin = np.zeros((self.nelemd, self.qsize, 72))
        qmax = np.zeros((self.nelemd, self.qsize, 72))
        self._ck(self._L.tse_get_qminmax(self._h, _p(qmin), _p(qmax)))
        return qmin, qmax

    def precompute_divdp(self):
        self._ck(self._L.tse_precompute_divdp(self._h))

    def euler_step(self, np1_qdp, n0_qdp, dt, DSSopt, rhs_multiplier):
        self._ck(self._L.tse_euler_step(self._h, np1_qdp, n0_qdp, dt, DSSopt, rhs_multiplier))

    def qdp_time_avg(self, rkstage, n0_qdp, np1_qdp):
        self._ck(self._L.tse_qdp_time_avg(self._h, rkstage, n0_qdp, np1_qdp))

    def vertical_remap(self, dt, np1, np1_qdp):
        self._ck(self._L.tse_vertical_remap(self._h, dt, np1, np1_qdp))

    def advance_hypervis_scalar(self, nt_qdp, dt2):
        """advance_hypervis_scalar_cuda (cuda_mod.F90:624-718): subcycled tracer hyperviscosity + limiter2d_zero; separate entry."""
        self._ck(self._L.tse_advance_hypervis_scalar(self._h, nt_qdp, dt2))

    def prim_advec_tracers_remap_rk2(self, dt, nstep):
        self._ck(self._L.tse_advec_tracers_remap_rk2(self._h, dt, nstep))

    # --- device-side test-case driver -----------------------------------------------------
    def dcmip_init(self, test_case):
        self._ck(self._L.tse_dcmip_init(self._h, test_case))

    def prim_run_subcycle(self, tstep, nstep):
        n = C.c_int(nstep)
        self._ck(self._L.tse_prim_run_subcycle(self._h, tstep, C.byref(n)))
        return n.value

    # --- diagnostics ----------------------------------------------------------------------
    def diag_mass(self, tl):
        out = np.zeros(self.qsize)
        self._ck(self._L.tse_diag_mass(self._h, tl, _p(out)))
        return out

    def diag_qminmax(self, tl):
        a, b = np.zeros(self.qsize), np.zeros(self.qsize)
        self._ck(self._L.tse_diag_qminmax(self._h, tl, _p(a), _p(b)))
        return a, b

    def diag_field_hash(self, tl):
        """Per-tracer fingerprint of Qdp(tl), independent of partition and element order (bit-for-bit checks across GPU counts)."""
        out = np.zeros(self.qsize, dtype=np.uint64)
        self._ck(self._L.tse_diag_field_hash(self._h, tl, out.ctypes.data_as(C.POINTER(C.c_ulonglong))))
        return out

    def synchronize(self):
        self._ck(self._L.tse_synchronize(self._h))

    def get_wind(self, vn0=None, dp=None):
        self._ck(self._L.tse_get_wind(self._h, _p(vn0), _stride(vn0), _p(dp), _stride(dp)))

    def mark(self, slot):
        self._ck(self._L.tse_mark(self._h, slot))

    def mark_elapsed_ms(self, a, b):
        return self._L.tse_mark_elapsed_ms(self._h, a, b)

    def timer_reset(self):
        self._ck(self._L.tse_timer_reset(self._h))

    @property
    def stage_launch_count(self):
        return self._L.tse_stage_launch_count(self._h)

    def timer_ms(self, name):
        return self._L.tse_timer_ms(self._h, name.encode())

    @property
    def launch_count(self):
        return self._L.tse_launch_count(self._h)

    @property
    def device_bytes(self):
        return self._L.tse_device_bytes(self._h)


def debug_limiter(ptens, sphweights, dpmass, minp, maxp):
    """limiter_optim_iter_full (prim_advection_mod.F90:976-1094) on n independent planes, through the device code the stage
    kernels use.  ptens/sphweights/dpmass: [n, 16]; minp/maxp: [n].  Returns (sphweights*ptens_limited, minp, maxp)."""
    L = cuda_lib()
    y = np.ascontiguousarray(ptens, dtype=np.float64).copy()
    sw = np.ascontiguousarray(sphweights, dtype=np.float64)
    dm = np.ascontiguousarray(dpmass, dtype=np.float64)
    mn = np.ascontiguousarray(minp, dtype=np.float64).copy()
    mx = np.ascontiguousarray(maxp, dtype=np.float64).copy()
    n = y.shape[0]
    assert y.shape == (n, 16) and sw.shape == (n, 16) and dm.shape == (n, 16) and mn.shape == (n,) and mx.shape == (n,)
    if L.tse_debug_limiter(n, _p(y), _p(sw), _p(dm), _p(mn), _p(mx)):
        raise TseError(L.tse_last_error().decode())
    return y, mn, mx
