"""Per-tracer mass drift after a few remap cycles (size-independent conservation check).  usage: mass_check.py ne cycles [qsize]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from transport_se_b200.mesh import Mesh, load_vcoord
from transport_se_b200.advection import TracerAdvection
ne, cycles = int(sys.argv[1]), int(sys.argv[2])
qsize = int(sys.argv[3]) if len(sys.argv) > 3 else 35
mesh = Mesh(ne); view = mesh.local_view(0, 1); hv = load_vcoord()
adv = TracerAdvection(mesh, view, hv, qsize=qsize, nu_q={8: 6e16, 30: 1e15}.get(ne, 1e13), device=0)
adv.dcmip_init(11)
m0 = adv.diag_mass(1)
tstep = {8: 400.0, 30: 300.0}.get(ne, 75.0 * 120 / ne)
nstep = 0
for c in range(cycles):
    nstep = adv.prim_run_subcycle(tstep, nstep)
    m = adv.diag_mass(1 if nstep % 2 == 0 else 2)
    d = np.abs(m - m0) / np.abs(m0)
    print('ne', ne, 'cycle', c, 'max drift %.3e at tracer %d; drift[0:6]' % (d.max(), int(d.argmax())), ' '.join('%.2e' % x for x in d[:6]), 'mass0[0]=%.12f' % m[0], flush=True)
adv.close()
