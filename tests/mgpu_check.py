"""Multi-GPU check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py [ne] [qsize] [test] [cycles]

Every rank advances its space-filling-curve chunk of the sphere (halo exchange over NCCL); rank 0 also advances the whole
sphere alone on its GPU.  The N-rank result must be BIT-FOR-BIT the single-rank result (BASELINE.json north_star), and the
order-independent mass diagnostic must agree bitwise too."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from transport_se_b200.mesh import Mesh, load_vcoord
from transport_se_b200.advection import TracerAdvection

NU_Q = {8: 6e16, 30: 1e15}
TSTEP = {8: 400.0, 30: 300.0}


def main():
    ne = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    qsize = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    test = int(sys.argv[3]) if len(sys.argv) > 3 else 11
    cycles = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mesh, hv = Mesh(ne), load_vcoord()
    tstep, nu_q = TSTEP.get(ne, 300.0), NU_Q.get(ne, 1e15)

    def run(view, with_comm):
        adv = TracerAdvection(mesh, view, hv, qsize=qsize, nu_q=nu_q, device=local)
        if with_comm:
            adv.comm_init(dist, rank, world)
        adv.dcmip_init(test)
        nstep = 0
        for _ in range(cycles):
            nstep = adv.prim_run_subcycle(tstep, nstep)
        # one more tracer step through the stage-by-stage entries, leaving a pending DSS for the d2h to resolve
        adv.set_derived()  # no-op
        tl = 1 if nstep % 2 == 0 else 2
        mass = adv.diag_mass(tl)
        out = np.zeros((view.nelemd, 2, qsize, 72, 16))
        adv.copy_qdp_d2h(out, tl)
        proj = np.zeros((view.nelemd, 72, 16))
        adv.get_derived(divdp_proj=proj)
        adv.synchronize()
        hb = adv.halo_bytes
        adv.close()
        return out[:, tl - 1].copy(), mass, proj, hb

    view = mesh.local_view(rank, world)
    q_loc, mass_loc, proj_loc, hb = run(view, True)
    parts = [None] * world
    dist.all_gather_object(parts, (view.gid, q_loc, proj_loc, mass_loc, hb))
    ok = True
    if rank == 0:
        q1, mass1, proj1, _ = run(mesh.local_view(0, 1), False)
        qn, pn = np.zeros_like(q1), np.zeros_like(proj1)
        for gid, q, p, m, b in parts:
            qn[gid] = q
            pn[gid] = p
            if not np.array_equal(m, mass1):
                ok = False
                print("mass differs from the single-rank run:", m, mass1)
        same_q, same_p = np.array_equal(qn, q1), np.array_equal(pn, proj1)
        print("mgpu_check ne=%d qsize=%d test=%d ranks=%d: Qdp bitwise %s, divdp_proj bitwise %s, max|dQ|=%.3e, halo bytes/rank %s"
              % (ne, qsize, test, world, same_q, same_p, np.max(np.abs(qn - q1)), [p[4] for p in parts]))
        ok = ok and same_q and same_p
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()
