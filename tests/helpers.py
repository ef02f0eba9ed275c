"""Shared test scaffolding: builds a case (mesh, rank view, vertical coordinate, oracle) and restates the
caller-side sequencing of prim_step (reference src/share/prim_driver_mod.F90:858-943) around the oracle so the
CUDA path can be fed the same inputs stage by stage."""
import functools

import numpy as np

from transport_se_b200.mesh import Mesh, load_vcoord

NU_Q = {8: 6e16, 30: 1e15, 120: 1e13}      # test/run_ne*_tests.sh
TSTEP = {8: 400.0, 30: 300.0, 120: 75.0}


@functools.lru_cache(maxsize=4)
def mesh_for(ne):
    return Mesh(ne)


def make_oracle(ne, qsize, test=11, nu_q=None):
    from oracle.oracle_lib import Oracle
    m = mesh_for(ne)
    v = m.local_view()
    hv = load_vcoord()
    nu = NU_Q.get(ne, 1e15) if nu_q is None else nu_q
    o = Oracle(m, v, hv, qsize=qsize, nu_q=nu)
    o.set_params(nu, 3, 8, test)
    o.prim_init2(test)
    return m, v, hv, o


def oracle_begin_step(o, test, tstep):
    """prim_step up to (not including) Prim_Advec_Tracers_remap: zero accumulators, derived%dp = dp3d(n0),
    prim_advance_exp (prim_advance_mod.F90:111-149)."""
    tl = o.tl
    o.eta_dot_dpdn[:] = 0.0
    o.vn0[:] = 0.0
    o.omega_p[:] = 0.0
    o.dp[:] = o.dp3d[:, tl["n0"] - 1]
    o.set_dcmip_fields(test, tl["np1"], tl["nstep"] * tstep)
    o.vn0 += o.v[:, tl["n0"] - 1] * o.dp[:, :, None, :]


def oracle_time_update(o):
    """TimeLevel_update('leapfrog') (time_mod.F90:111-140)"""
    tl = o.tl
    o.set_tl(tl["n0"], tl["np1"], tl["nm1"], tl["nstep"] + 1)


def relerr(a, b, axis=None):
    """max-norm relative error of a against b"""
    den = np.max(np.abs(b), axis=axis)
    den = np.where(den == 0, 1.0, den)
    return np.max(np.abs(a - b), axis=axis) / den


def per_tracer_relerr(qa, qb):
    """qa,qb: [e, q, k, 16] -> [q]"""
    return relerr(qa, qb, axis=(0, 2, 3))
