"""Host-side cubed-sphere mesh (ctypes view of libtse_host.so).

This is the caller side of the hot path: it produces the inputs the reference's
Fortran host hands to prim_advection_mod -- element_t metric terms
(reference src/share/element_mod.F90:112-221), derivative_t%Dvv
(src/share/derivative_mod.F90:451-486), EdgeDescriptor_t put/get maps
(src/share/edge_mod.F90:31-43), the SFC partition
(src/share/spacecurve_mod.F90:1218-1273) and the hybrid vertical coordinate
(src/share/hybvcoord_mod.F90:18-30).  Pure host code; no CUDA.
"""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def host_lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libtse_host.so")
        if not os.path.exists(path):
            from . import _build
            _build.build_host()
        lib = C.CDLL(path)
        lib.tse_mesh_create.restype = C.c_void_p
        lib.tse_mesh_create.argtypes = [C.c_int]
        lib.tse_mesh_destroy.argtypes = [C.c_void_p]
        lib.tse_mesh_nelem.argtypes = [C.c_void_p]
        lib.tse_mesh_alpha.restype = C.c_double
        lib.tse_mesh_alpha.argtypes = [C.c_void_p]
        lib.tse_mesh_get.argtypes = [C.c_void_p] + [_dp] * 8 + [_ip] * 5
        lib.tse_mesh_sfc_partition.argtypes = [C.c_void_p, C.c_int, _ip]
        lib.tse_local_view_create.restype = C.c_void_p
        lib.tse_local_view_create.argtypes = [C.c_void_p, _ip, C.c_int, C.c_int]
        lib.tse_local_view_destroy.argtypes = [C.c_void_p]
        lib.tse_local_view_sizes.argtypes = [C.c_void_p, _ip, _ip, _ip]
        lib.tse_local_view_get.argtypes = [C.c_void_p] + [_ip] * 7
        lib.tse_gll.argtypes = [_dp] * 4
        _LIB = lib
    return _LIB


def gll():
    pts, wts, dvv, mp = (np.zeros(4), np.zeros(4), np.zeros(16), np.zeros(16))
    host_lib().tse_gll(_p(pts), _p(wts), _p(dvv), _p(mp))
    return pts, wts, dvv, mp


def load_vcoord(name="acme72"):
    """hvcoord_t for the shipped 72-level ACME tables (reference test/vcoord/acme-72{i,m}.ascii)."""
    with open(os.path.join(_HERE, "data", "%s_vcoord.json" % name)) as f:
        d = json.load(f)
    return {k: np.asarray(d[k], dtype=np.float64) for k in ("hyai", "hybi", "hyam", "hybm")}


class LocalView:
    """Per-rank element list + EdgeDescriptor_t maps + exchange cycles."""

    def __init__(self, mesh, owner, rank, nranks):
        lib = host_lib()
        owner = np.ascontiguousarray(owner, dtype=np.int32)
        h = lib.tse_local_view_create(mesh._h, _p(owner, C.c_int), rank, nranks)
        n, nbuf, ncyc = C.c_int(), C.c_int(), C.c_int()
        lib.tse_local_view_sizes(h, C.byref(n), C.byref(nbuf), C.byref(ncyc))
        self.rank, self.nranks = rank, nranks
        self.nelemd, self.nbuf, self.ncycles = n.value, nbuf.value, ncyc.value
        self.gid = np.zeros(self.nelemd, np.int32)
        self.putmap = np.zeros((self.nelemd, 8), np.int32)
        self.getmap = np.zeros((self.nelemd, 8), np.int32)
        self.reverse = np.zeros((self.nelemd, 8), np.int32)
        self.cyc_rank = np.zeros(self.ncycles, np.int32)
        self.cyc_ptr = np.zeros(self.ncycles, np.int32)
        self.cyc_len = np.zeros(self.ncycles, np.int32)
        lib.tse_local_view_get(h, _p(self.gid, C.c_int), _p(self.putmap, C.c_int), _p(self.getmap, C.c_int),
                               _p(self.reverse, C.c_int), _p(self.cyc_rank, C.c_int), _p(self.cyc_ptr, C.c_int),
                               _p(self.cyc_len, C.c_int))
        lib.tse_local_view_destroy(h)


class Mesh:
    """Uniform cubed sphere, np=4, equi-angular map (reference cube_mod.F90)."""

    def __init__(self, ne):
        lib = host_lib()
        self.ne = ne
        self._h = lib.tse_mesh_create(ne)
        if not self._h:
            raise RuntimeError("mesh creation failed for ne=%d" % ne)
        n = self.nelem = lib.tse_mesh_nelem(self._h)
        self.alpha = lib.tse_mesh_alpha(self._h)
        self.lat = np.zeros((n, 16)); self.lon = np.zeros((n, 16))
        self.D = np.zeros((n, 16, 4)); self.Dinv = np.zeros((n, 16, 4))
        self.metdet = np.zeros((n, 16)); self.rmetdet = np.zeros((n, 16))
        self.spheremp = np.zeros((n, 16)); self.rspheremp = np.zeros((n, 16))
        self.nbr = np.zeros((n, 8), np.int32); self.nbr_dir = np.zeros((n, 8), np.int32)
        self.rev = np.zeros((n, 8), np.int32); self.sfc = np.zeros(n, np.int32)
        self.face_ie_je = np.zeros((n, 3), np.int32)
        lib.tse_mesh_get(self._h, _p(self.lat), _p(self.lon), _p(self.D), _p(self.Dinv), _p(self.metdet), _p(self.rmetdet),
                         _p(self.spheremp), _p(self.rspheremp), _p(self.nbr, C.c_int), _p(self.nbr_dir, C.c_int),
                         _p(self.rev, C.c_int), _p(self.sfc, C.c_int), _p(self.face_ie_je, C.c_int))
        self.pts, self.wts, self.dvv, self.mp = gll()

    def __del__(self):
        try:
            if self._h:
                host_lib().tse_mesh_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def sfc_partition(self, nparts):
        owner = np.zeros(self.nelem, np.int32)
        host_lib().tse_mesh_sfc_partition(self._h, nparts, _p(owner, C.c_int))
        return owner

    def local_view(self, rank=0, nranks=1, owner=None):
        if owner is None:
            owner = self.sfc_partition(nranks)
        return LocalView(self, owner, rank, nranks)
