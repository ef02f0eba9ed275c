"""One rank's share of an N-GPU run on a single GPU, without the exchange (TSE_PROFILE_NO_EXCHANGE=1: results are meaningless, the
kernels, launch splits and packs are those of the real run).  usage: rank_profile.py ne nranks rank [cycles]   (run under ncu)"""
import os, sys, time
os.environ["TSE_PROFILE_NO_EXCHANGE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transport_se_b200.mesh import Mesh, load_vcoord
from transport_se_b200.advection import TracerAdvection
ne, nranks, rank = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cycles = int(sys.argv[4]) if len(sys.argv) > 4 else 2
m = Mesh(ne)
adv = TracerAdvection(m, m.local_view(rank, nranks), load_vcoord(), qsize=35, nu_q={30: 1e15, 120: 1e13}.get(ne, 1e13))
adv.dcmip_init(11)
tstep = {30: 300.0, 120: 75.0}.get(ne, 75.0)
nstep = adv.prim_run_subcycle(tstep, 0)
adv.synchronize()
adv.mark(0)
for _ in range(cycles):
    nstep = adv.prim_run_subcycle(tstep, nstep)
adv.mark(1)
adv.synchronize()
print("rank %d of %d, ne%d: %.3f ms per tracer step (no exchange)" % (rank, nranks, ne, adv.mark_elapsed_ms(0, 1) / (3 * cycles)))
