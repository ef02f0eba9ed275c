"""Counting experiment (build with tools/build_variant.sh limstats -DTSE_EXP_LIMSTATS, run with TSE_CUDA_LIB pointing at it):
how many warps reach the limiter, how many enter the sweeps, and how many lanes are active there (ne120 perf case, one cycle)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transport_se_b200.advection import TracerAdvection, cuda_lib  # noqa: E402
from transport_se_b200.mesh import Mesh, load_vcoord  # noqa: E402

ne, qsize, test = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
tstep = {8: 400.0, 30: 300.0, 120: 75.0}[ne]
nu_q = {8: 6e16, 30: 1e15, 120: 1e13}[ne]
m = Mesh(ne)
adv = TracerAdvection(m, m.local_view(0, 1), load_vcoord(), qsize=qsize, nu_q=nu_q, device=0)
adv.dcmip_init(test)
L = cuda_lib()
out = (C.c_ulonglong * 4)()
nstep = 0
for cyc in range(3):
    L.tse_exp_limiter_stats(out, 1)
    nstep = adv.prim_run_subcycle(tstep, nstep)
    L.tse_exp_limiter_stats(out, 0)
    w, l, ws, ls = [int(x) for x in out]
    print(f"cycle {cyc}: warps {w} lanes {l}; warps entering sweeps {ws} ({ws / w:.3f}); planes needing sweeps {ls} ({ls / l:.3f}); "
          f"active lanes per entering warp {ls / max(ws, 1):.1f} of {l / w:.1f}")
adv.close()
