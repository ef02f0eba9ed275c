"""Pins the CPU oracle against the reference's published end-of-run error norms.

Runs the full DCMIP 1-1 (12 days) / 1-2 (1 day) verification configuration of
test/run_ne8_tests.sh (ne=8, tstep=400, nu_q=6e16, qsize=4, rsplit=3, limiter 8, 72 ACME levels)
through the oracle and writes the NCL norms next to the README values (reference README:94-96).

usage: python tests/golden/make_oracle_norms.py <ne> <test 11|12> [tstep nu_q]
Takes minutes (ne8) to hours (ne30) of CPU; result committed as tests/golden/oracle_norms_*.json.
"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from transport_se_b200.mesh import Mesh, load_vcoord
from transport_se_b200.diagnostics import dcmip_error_norms
from oracle.oracle_lib import Oracle

CFG = {8: (400.0, 6e16), 30: (300.0, 1e15), 120: (75.0, 1e13)}
README = {(8, 11): dict(L1=0.578151, L2=0.865526, Linf=0.883168, q_max=0.187204, q_min=-3.207090e-13),
          (8, 12): dict(L1=0.307665, L2=0.622099, Linf=0.839133, q_max=0.813105, q_min=-9.385639e-06),
          (30, 11): dict(L1=0.490013, L2=0.789052, Linf=0.918454, q_max=0.445141, q_min=-3.559994e-11),
          (30, 12): dict(L1=0.121783, L2=0.361005, Linf=1.092784, q_max=0.836177, q_min=-3.671997e-05)}

ne, test = int(sys.argv[1]), int(sys.argv[2])
tstep, nu_q = CFG[ne]
if len(sys.argv) > 4:
    tstep, nu_q = float(sys.argv[3]), float(sys.argv[4])
ndays = 12 if test == 11 else 1
nsteps = int(round(ndays * 86400 / tstep))
m = Mesh(ne); v = m.local_view(); hv = load_vcoord()
o = Oracle(m, v, hv, qsize=4, nu_q=nu_q)
o.set_params(nu_q, 3, 8, test)
o.prim_init2(test)
tracer = 0 if test == 11 else 1   # NCL: Q for 1-1, Q2 for 1-2
q_i = o.Q[:, tracer].copy()
z_mid = o.phi[0, :, 0] / 9.80616
mass0 = [(o.Qdp[:, 0, q] * m.spheremp[:, None, :]).sum() for q in range(4)]
t0 = time.time()
assert nsteps % 3 == 0
for i in range(nsteps // 3):
    bad = o.prim_run_subcycle(tstep)
    assert bad == 0
q_f = o.Q[:, tracer].copy()
n0, np1 = o.qdp_levels()
mass1 = [(o.Qdp[:, n0 - 1, q] * m.spheremp[:, None, :]).sum() for q in range(4)]
res = dcmip_error_norms(m, q_i, q_f, z_mid)
out = dict(ne=ne, test=test, tstep=tstep, nu_q=nu_q, nsteps=nsteps, oracle=res, readme=README.get((ne, test)),
           mass_rel_drift=[float((a - b) / b) if b != 0 else 0.0 for a, b in zip(mass1, mass0)], wall_s=time.time() - t0)
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_norms_ne%d_dcmip%d.json" % (ne, test)), "w"), indent=1)
