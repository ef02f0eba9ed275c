"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): per-tracer fields within 1e-12 relative max-norm after one step.
The tests hold every stage to TOL = 1e-12 and print the achieved error."""
import numpy as np
import pytest

from helpers import make_oracle, oracle_begin_step, oracle_time_update, per_tracer_relerr, relerr, TSTEP

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _device(m, v, hv, o, qsize, nu_q):
    from transport_se_b200.advection import TracerAdvection
    adv = TracerAdvection(m, v, hv, qsize=qsize, nu_q=nu_q)
    adv.copy_qdp_h2d(o.Qdp, 1)
    adv.copy_qdp_h2d(o.Qdp, 2)
    return adv


def _fetch(adv, o, tl):
    out = np.zeros_like(o.Qdp)
    adv.copy_qdp_d2h(out, tl)
    return out[:, tl - 1]


@pytest.mark.parametrize("test,qsize", [(11, 4), (12, 5)])
def test_stage_by_stage(built, test, qsize):
    from oracle.oracle_lib import DSSeta, DSSomega, DSSdiv_vdp_ave
    ne = 8
    tstep = TSTEP[ne]
    m, v, hv, o = make_oracle(ne, qsize, test)
    adv = _device(m, v, hv, o, qsize, 6e16)
    # round trip of the layout conversion
    assert np.array_equal(_fetch(adv, o, 1), o.Qdp[:, 0])
    for step in range(3):
        oracle_begin_step(o, test, tstep)
        adv.set_derived(o.vn0, o.dp, o.eta_dot_dpdn, o.omega_p)
        n0, np1 = o.qdp_levels()
        o.precompute_divdp()
        adv.precompute_divdp()
        dd = np.zeros_like(o.divdp)
        adv.get_derived(divdp=dd)
        assert relerr(dd, o.divdp) < TOL
        for rhs, dss, a, b in ((0, DSSdiv_vdp_ave, np1, n0), (1, DSSeta, np1, np1), (2, DSSomega, np1, np1)):
            o.euler_step(a, b, tstep / 2, dss, rhs)
            adv.euler_step(a, b, tstep / 2, dss, rhs)
            err = per_tracer_relerr(_fetch(adv, o, np1), o.Qdp[:, np1 - 1])
            print("step", step, "stage", rhs + 1, "relerr per tracer", err)
            assert err.max() < TOL
            qmin, qmax = adv.get_qminmax()
            assert relerr(qmin, o.qmin) < TOL and relerr(qmax, o.qmax) < TOL
            proj, eta, om = np.zeros_like(o.divdp_proj), np.zeros_like(o.eta_dot_dpdn), np.zeros_like(o.omega_p)
            adv.get_derived(divdp_proj=proj, eta_dot_dpdn=eta, omega_p=om)
            assert relerr(proj, o.divdp_proj) < TOL
            assert relerr(eta[:, :72], o.eta_dot_dpdn[:, :72]) < TOL
        o.qdp_time_avg(3, n0, np1)
        adv.qdp_time_avg(3, n0, np1)
        err = per_tracer_relerr(_fetch(adv, o, np1), o.Qdp[:, np1 - 1])
        print("step", step, "time_avg relerr", err)
        assert err.max() < TOL
        if step < 2:
            oracle_time_update(o)
    # vertical remap at the end of the rsplit=3 cycle
    n0, np1 = o.qdp_levels()
    tl = o.tl
    assert o.vertical_remap(3 * tstep, tl["np1"], np1) == 0
    adv.vertical_remap(3 * tstep, tl["np1"], np1)
    adv.synchronize()
    err = per_tracer_relerr(_fetch(adv, o, np1), o.Qdp[:, np1 - 1])
    print("remap relerr", err)
    assert err.max() < TOL
    dp3d, ps = np.zeros_like(o.dp), np.zeros((o.nelem, 16))
    adv.get_dp3d_ps(dp3d, ps)
    assert relerr(dp3d, o.dp3d[:, tl["np1"] - 1]) < TOL and relerr(ps, o.ps_v[:, tl["np1"] - 1]) < TOL
    adv.close()


def test_fused_two_cycles(built):
    """Prim_Advec_Tracers_remap_rk2 + vertical_remap through the fused entry, 2 remap cycles (6 steps)."""
    ne, qsize, test = 8, 4, 11
    tstep = TSTEP[ne]
    m, v, hv, o = make_oracle(ne, qsize, test)
    adv = _device(m, v, hv, o, qsize, 6e16)
    mass0 = (o.Qdp[:, 0] * m.spheremp[:, None, None, :]).sum(axis=(0, 2, 3))
    for cyc in range(2):
        for r in range(3):
            if r > 0:
                oracle_time_update(o)
            oracle_begin_step(o, test, tstep)
            adv.set_derived(o.vn0, o.dp, o.eta_dot_dpdn, o.omega_p)
            o.advec_tracers_remap_rk2(tstep)
            adv.prim_advec_tracers_remap_rk2(tstep, o.tl["nstep"])
        n0, np1 = o.qdp_levels()
        assert o.vertical_remap(3 * tstep, o.tl["np1"], np1) == 0
        adv.vertical_remap(3 * tstep, o.tl["np1"], np1)
        err = per_tracer_relerr(_fetch(adv, o, np1), o.Qdp[:, np1 - 1])
        print("cycle", cyc, "relerr", err)
        assert err.max() < 5e-12  # accumulated over 3 steps + remap
        oracle_time_update(o)
    g = _fetch(adv, o, np1)
    mass1 = (g * m.spheremp[:, None, None, :]).sum(axis=(0, 2, 3))
    assert np.max(np.abs(mass1 - mass0) / mass0) < 1e-13  # tracer mass conserved to roundoff
    adv.close()


def test_constant_tracer_consistency(built):
    """Q == 1 stays Qdp == dp through euler_step when nu_q = 0 (consistency note prim_advection_mod.F90:23-38)."""
    from oracle.oracle_lib import DSSdiv_vdp_ave
    ne, qsize, test = 8, 4, 11
    m, v, hv, o = make_oracle(ne, qsize, test, nu_q=0.0)
    oracle_begin_step(o, test, TSTEP[ne])
    o.Qdp[:] = o.dp[:, None, None]
    adv = _device(m, v, hv, o, qsize, 0.0)
    adv.set_derived(o.vn0, o.dp, o.eta_dot_dpdn, o.omega_p)
    adv.precompute_divdp()
    dt = TSTEP[ne] / 2
    adv.euler_step(2, 1, dt, DSSdiv_vdp_ave, 0)
    g = _fetch(adv, o, 2)
    o.precompute_divdp()
    dd = np.zeros_like(o.divdp)
    adv.get_derived(divdp=dd)
    # Qdp(np1) must equal DSS(spheremp*(dp - dt*divdp))*rspheremp
    ref = (o.dp - dt * dd) * m.spheremp[:, None, :]
    ref = np.ascontiguousarray(ref)
    o.dss(ref)
    ref *= m.rspheremp[:, None, :]
    assert relerr(g, np.broadcast_to(ref[:, None], g.shape)) < 1e-12
    adv.close()
