"""Multi-GPU check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py [ne] [qsize] [test] [cycles] [limiter_option]

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
from transport_se_b200.advection import TracerAdvection, DSSeta, DSSdiv_vdp_ave, DSSno_var

NU_Q = {8: 6e16, 30: 1e15}
TSTEP = {8: 400.0, 30: 300.0}


def main():
    ne = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    qsize = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    test = int(sys.argv[3]) if len(sys.argv) > 3 else 11
    cycles = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    limiter_option = int(sys.argv[5]) if len(sys.argv) > 5 else 8
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mesh, hv = Mesh(ne), load_vcoord()
    tstep, nu_q = TSTEP.get(ne, 300.0), NU_Q.get(ne, 1e15)

    def run(view, with_comm):
        adv = TracerAdvection(mesh, view, hv, qsize=qsize, nu_q=nu_q, device=local, limiter_option=limiter_option)
        if with_comm:
            adv.comm_init(dist, rank, world)
        adv.dcmip_init(test)
        nstep = 0
        for _ in range(cycles):
            nstep = adv.prim_run_subcycle(tstep, nstep)
        tl = 1 if nstep % 2 == 0 else 2
        mass = adv.diag_mass(tl)
        fh = adv.diag_field_hash(tl)
        out = np.zeros((view.nelemd, 2, qsize, 72, 16))
        adv.copy_qdp_d2h(out, tl)
        proj = np.zeros((view.nelemd, 72, 16))
        adv.get_derived(divdp_proj=proj)
        # One more tracer step through the stage-by-stage entries WITHOUT the time average / remap: time level np1 is left
        # pending (pre-DSS values + ghosts of the last exchange), so the diagnostics and the d2h copy below go through the paths
        # that read the ghost array outside the stage kernels (DssView::load in k_q_minmax / k_field_hash, the halo of OP_MASS,
        # OP_RESOLVE with ghosts) and through single euler_step calls, the last one with DSS_NO_VAR (tracer-only exchange).
        np1 = 3 - tl
        adv.precompute_divdp()
        adv.euler_step(np1, tl, tstep / 2, DSSdiv_vdp_ave, 0)
        adv.euler_step(np1, np1, tstep / 2, DSSeta, 1)
        adv.euler_step(np1, np1, tstep / 2, DSSno_var, 2)
        adv.advance_hypervis_scalar(np1, tstep)   # the separate hyperviscosity entry: two more exchanges through the same paths
        mass_p = adv.diag_mass(np1)
        qmn_p, qmx_p = adv.diag_qminmax(np1)
        fh_p = adv.diag_field_hash(np1)
        adv.copy_qdp_d2h(out, np1)
        adv.synchronize()
        hb = adv.halo_bytes
        adv.close()
        return out[:, tl - 1].copy(), (mass, mass_p, qmn_p, qmx_p, fh, fh_p), proj, hb, out[:, np1 - 1].copy()

    view = mesh.local_view(rank, world)
    q_loc, scal_loc, proj_loc, hb, qp_loc = run(view, True)
    parts = [None] * world
    dist.all_gather_object(parts, (view.gid, q_loc, proj_loc, scal_loc, hb, qp_loc))
    ok = True
    if rank == 0:
        q1, scal1, proj1, _, qp1 = run(mesh.local_view(0, 1), False)
        qn, pn, qpn = np.zeros_like(q1), np.zeros_like(proj1), np.zeros_like(qp1)
        names = ("mass", "mass(pending level)", "qmin(pending level)", "qmax(pending level)", "field hash", "field hash(pending level)")
        for gid, q, p, sc, b, qp in parts:
            qn[gid] = q
            pn[gid] = p
            qpn[gid] = qp
            for nm, x, y in zip(names, sc, scal1):
                if not np.array_equal(x, y):
                    ok = False
                    print("%s differs from the single-rank run:" % nm, x, y)
        same_q, same_p, same_qp = np.array_equal(qn, q1), np.array_equal(pn, proj1), np.array_equal(qpn, qp1)
        print("mgpu_check ne=%d qsize=%d test=%d limiter=%d ranks=%d: Qdp bitwise %s, divdp_proj bitwise %s, Qdp after 3 more stages (resolved "
              "from a pending level) bitwise %s, diagnostics bitwise %s, max|dQ|=%.3e, halo bytes/rank %s"
              % (ne, qsize, test, limiter_option, world, same_q, same_p, same_qp, ok, np.max(np.abs(qn - q1)), [p[4] for p in parts]))
        ok = ok and same_q and same_p and same_qp
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag[0]) == 1 else 1)


if __name__ == "__main__":
    main()
