"""ctypes view of the CPU oracle (oracle/oracle.cpp) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  Nothing under transport_se_b200/ does.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

DSSeta, DSSomega, DSSdiv_vdp_ave, DSSno_var = 1, 2, 3, -1


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int] + [_dp] * 8 + [_ip] * 3 + [C.c_int] + [_dp] * 4 + [C.c_double, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_set_num_threads.restype = C.c_int
        L.orc_field.restype = _dp
        L.orc_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_longlong)]
        L.orc_get_tl.argtypes = [C.c_void_p, _ip]
        L.orc_set_tl.argtypes = [C.c_void_p, _ip]
        L.orc_set_params.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]
        L.orc_prim_init2.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_dcmip_fields.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_prim_step.argtypes = [C.c_void_p, C.c_double]
        L.orc_prim_run_subcycle.argtypes = [C.c_void_p, C.c_double]
        L.orc_prim_run_subcycle.restype = C.c_int
        L.orc_precompute_divdp.argtypes = [C.c_void_p]
        L.orc_euler_step.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
        L.orc_qdp_time_avg.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_advec_tracers_remap_rk2.argtypes = [C.c_void_p, C.c_double]
        L.orc_vertical_remap.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int]
        L.orc_vertical_remap.restype = C.c_int
        L.orc_neighbor_minmax.argtypes = [C.c_void_p]
        L.orc_advance_hypervis_scalar.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
        L.orc_dss.argtypes = [C.c_void_p, _dp, C.c_int]
        L.orc_global_integral.restype = C.c_double
        L.orc_global_integral.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_divergence_sphere.argtypes = [_dp] * 6
        L.orc_gradient_sphere.argtypes = [_dp] * 4
        L.orc_divergence_sphere_wk.argtypes = [_dp] * 5
        L.orc_laplace_sphere_wk.argtypes = [_dp] * 5
        L.orc_limiter_optim_iter_full.argtypes = [_dp] * 5
        L.orc_remap_q_ppm.argtypes = [_dp, C.c_int, C.c_int, C.c_int, _dp, _dp]
        L.orc_dcmip_point.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp]
        _LIB = L
    return _LIB


def set_num_threads(n):
    """OpenMP threads of the oracle's element loops; returns the count in effect (n <= 0: just query)."""
    return lib().orc_set_num_threads(int(n))


class Oracle:
    """One rank of the reference path: state in the reference's own element-major layout.

    Fields (numpy views into the oracle's memory, C order, (i,j) fastest = last axis of 16):
      Qdp[e,tl2,q,k,16]  v[e,tl3,k,c,16]  dp3d[e,tl3,k,16]  ps_v[e,tl3,16]  Q[e,q,k,16]
      vn0[e,k,c,16]  dp,divdp,divdp_proj,omega_p,phi[e,k,16]  eta_dot_dpdn[e,k+1,16]  qmin,qmax[e,q,k]
    """

    def __init__(self, mesh, view, vcoord, qsize, nlev=72, nu_q=0.0, rsplit=3):
        L = lib()
        g = view.gid
        self.nelem, self.qsize, self.nlev = len(g), qsize, nlev
        self.mesh, self.view = mesh, view
        c = np.ascontiguousarray
        self._keep = [c(mesh.dvv), c(mesh.spheremp[g]), c(mesh.rspheremp[g]), c(mesh.metdet[g]), c(mesh.rmetdet[g]),
                      c(mesh.Dinv[g]), c(mesh.lat[g]), c(mesh.lon[g])]
        ints = [c(view.putmap), c(view.getmap), c(view.reverse)]
        hv = [c(vcoord[k]) for k in ("hyai", "hybi", "hyam", "hybm")]
        self._h = L.orc_create(self.nelem, qsize, nlev, *[_p(a) for a in self._keep], *[_p(a, C.c_int) for a in ints],
                               view.nbuf, *[_p(a) for a in hv], nu_q, rsplit)
        n, q, k = self.nelem, qsize, nlev
        shapes = dict(Qdp=(n, 2, q, k, 16), v=(n, 3, k, 2, 16), dp3d=(n, 3, k, 16), ps_v=(n, 3, 16), Q=(n, q, k, 16),
                      vn0=(n, k, 2, 16), dp=(n, k, 16), divdp=(n, k, 16), divdp_proj=(n, k, 16), omega_p=(n, k, 16),
                      phi=(n, k, 16), eta_dot_dpdn=(n, k + 1, 16), qmin=(n, q, k), qmax=(n, q, k))
        for name, shp in shapes.items():
            cnt = C.c_longlong()
            ptr = L.orc_field(self._h, name.encode(), C.byref(cnt))
            assert cnt.value == int(np.prod(shp)), (name, cnt.value, shp)
            setattr(self, name, np.ctypeslib.as_array(ptr, shape=shp))

    def __del__(self):
        try:
            if self._h:
                lib().orc_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # time levels (1-based like the reference)
    @property
    def tl(self):
        a = np.zeros(4, np.int32)
        lib().orc_get_tl(self._h, _p(a, C.c_int))
        return dict(nm1=int(a[0]), n0=int(a[1]), np1=int(a[2]), nstep=int(a[3]))

    def set_tl(self, nm1, n0, np1, nstep):
        a = np.array([nm1, n0, np1, nstep], np.int32)
        lib().orc_set_tl(self._h, _p(a, C.c_int))

    def qdp_levels(self):
        """TimeLevel_Qdp (time_mod.F90:85-109) -> (n0_qdp, np1_qdp), 1-based."""
        ns = self.tl["nstep"]
        return (1, 2) if ns % 2 == 0 else (2, 1)

    def set_params(self, nu_q, rsplit=3, limiter_option=8, test_case=11):
        lib().orc_set_params(self._h, nu_q, rsplit, limiter_option, test_case)

    def prim_init2(self, test):
        lib().orc_prim_init2(self._h, test)

    def set_dcmip_fields(self, test, tlv, time):
        lib().orc_set_dcmip_fields(self._h, test, tlv, time)

    def prim_step(self, dt):
        lib().orc_prim_step(self._h, dt)

    def prim_run_subcycle(self, dt):
        return lib().orc_prim_run_subcycle(self._h, dt)

    def precompute_divdp(self):
        lib().orc_precompute_divdp(self._h)

    def euler_step(self, np1_qdp, n0_qdp, dt, DSSopt, rhs_multiplier):
        lib().orc_euler_step(self._h, np1_qdp, n0_qdp, dt, DSSopt, rhs_multiplier)

    def qdp_time_avg(self, rkstage, n0_qdp, np1_qdp):
        lib().orc_qdp_time_avg(self._h, rkstage, n0_qdp, np1_qdp)

    def advec_tracers_remap_rk2(self, dt):
        lib().orc_advec_tracers_remap_rk2(self._h, dt)

    def vertical_remap(self, dt, np1, np1_qdp):
        return lib().orc_vertical_remap(self._h, dt, np1, np1_qdp)

    def neighbor_minmax(self):
        lib().orc_neighbor_minmax(self._h)

    def advance_hypervis_scalar(self, nt_qdp, dt2, hypervis_subcycle_q=1):
        """cuda_mod.F90:624-718 (not on the reference's CPU path; see the note in oracle.cpp)"""
        lib().orc_advance_hypervis_scalar(self._h, nt_qdp, dt2, hypervis_subcycle_q)

    def dss(self, field):
        assert field.flags.c_contiguous and field.shape[0] == self.nelem and field.shape[-1] == 16
        lib().orc_dss(self._h, _p(field), int(field.size // (self.nelem * 16)))

    def global_integral(self, h):
        h = np.ascontiguousarray(h)
        mp = np.ascontiguousarray(self.mesh.mp)
        return lib().orc_global_integral(self._h, _p(h), _p(mp))
