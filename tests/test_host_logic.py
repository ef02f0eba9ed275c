"""Host-side logic on the CPU: mesh topology, space-filling-curve partition, rank views, the C-ABI surface, and the
N>1 exchange schedule exercised with a world_size-2 gloo run (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import mesh_for

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ne", [4, 8, 30])
def test_topology(ne):
    m = mesh_for(ne)
    assert m.nelem == 6 * ne * ne
    # 8 neighbours everywhere except the 24 elements at the cube corners, which have 7 (cube_mod.F90:1638-1719)
    nn = (m.nbr >= 0).sum(axis=1)
    assert (nn == 7).sum() == 24 and (nn == 8).sum() == m.nelem - 24
    # neighbour relation is symmetric and the facing direction points back
    for e in range(0, m.nelem, 5):
        for d in range(8):
            b = m.nbr[e, d]
            if b >= 0:
                assert m.nbr[b, m.nbr_dir[e, d]] == e
    # reverse flags agree on both sides of an edge
    for e in range(0, m.nelem, 3):
        for d in range(4):
            b, bd = m.nbr[e, d], m.nbr_dir[e, d]
            assert m.rev[e, d] == m.rev[b, bd]


@pytest.mark.parametrize("ne", [4, 8, 30, 120])
def test_sfc_is_a_curve(ne):
    m = mesh_for(ne) if ne != 120 else __import__("transport_se_b200.mesh", fromlist=["Mesh"]).Mesh(ne)
    assert sorted(m.sfc.tolist()) == list(range(m.nelem))  # bijection onto 0..nelem-1
    # consecutive elements along the curve are edge neighbours (a continuous curve across the six faces)
    order = np.argsort(m.sfc)
    for a, b in zip(order[:-1], order[1:]):
        assert b in m.nbr[a, :4]


@pytest.mark.parametrize("ne,nparts", [(8, 2), (8, 3), (30, 7), (30, 8)])
def test_partition_chunks(ne, nparts):
    m = mesh_for(ne)
    owner = m.sfc_partition(nparts)
    counts = np.bincount(owner, minlength=nparts)
    base, extra = divmod(m.nelem, nparts)
    assert sorted(counts.tolist()) == sorted([base + 1] * extra + [base] * (nparts - extra))  # spacecurve_mod.F90:1235-1263
    # every rank owns one contiguous piece of the curve
    order = np.argsort(m.sfc)
    o = owner[order]
    assert (np.diff(o) >= 0).all()


def test_rank_views_are_consistent():
    m = mesh_for(8)
    nr = 4
    owner = m.sfc_partition(nr)
    views = [m.local_view(r, nr, owner) for r in range(nr)]
    assert sum(v.nelemd for v in views) == m.nelem
    for r, v in enumerate(views):
        assert (np.diff(v.gid) > 0).all()  # local order = ascending global id (metagraph_mod.F90:317-323)
        for c in range(v.ncycles):
            peer = views[v.cyc_rank[c]]
            back = list(peer.cyc_rank).index(r)
            assert peer.cyc_len[back] == v.cyc_len[c]  # both sides agree on the slab size


def _dss_worker(rank, world, port, ne, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    import torch
    m = mesh_for(ne)
    v = m.local_view(rank, world)
    rng = np.random.default_rng(5)
    field = rng.random((m.nelem, 3, 16))  # same global field on every rank
    loc = field[v.gid] * m.spheremp[v.gid][:, None, :]
    nl = loc.shape[1]
    buf = np.zeros((v.nbuf, nl))
    edge_nodes = {2: [0, 1, 2, 3], 1: [3, 7, 11, 15], 3: [12, 13, 14, 15], 0: [0, 4, 8, 12]}
    corner = {4: 0, 5: 3, 6: 12, 7: 15}
    # edgeVpack (edge_mod.F90:366-511)
    for le in range(v.nelemd):
        for d, nodes in edge_nodes.items():
            pm = v.putmap[le, d]
            for i, nd in enumerate(nodes):
                buf[pm + (3 - i if v.reverse[le, d] else i)] = loc[le, :, nd]
        for d, nd in corner.items():
            if v.putmap[le, d] >= 0:
                buf[v.putmap[le, d]] = loc[le, :, nd]
    # bndry_exchangeV (bndry_mod.F90:74-112): one message per neighbour rank
    recv = buf.copy()
    reqs, keep = [], []
    for c in range(v.ncycles):
        s = torch.from_numpy(buf[v.cyc_ptr[c]:v.cyc_ptr[c] + v.cyc_len[c]].copy())
        r_ = torch.zeros_like(s)
        keep.append((c, r_))
        reqs.append(dist.isend(s, int(v.cyc_rank[c])))
        reqs.append(dist.irecv(r_, int(v.cyc_rank[c])))
    for rq in reqs:
        rq.wait()
    for c, r_ in keep:
        recv[v.cyc_ptr[c]:v.cyc_ptr[c] + v.cyc_len[c]] = r_.numpy()
    # edgeVunpack (edge_mod.F90:648-742), reference order S, E, N, W, then SW, SE, NE, NW
    out = loc.copy()
    for le in range(v.nelemd):
        for d in (2, 1, 3, 0):
            gm = v.getmap[le, d]
            for i, nd in enumerate(edge_nodes[d]):
                out[le, :, nd] += recv[gm + i]
        for d in (4, 5, 7, 6):
            if v.getmap[le, d] >= 0:
                out[le, :, corner[d]] += recv[v.getmap[le, d]]
    out *= m.rspheremp[v.gid][:, None, :]
    q.put((rank, v.gid.copy(), out))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_dss_matches_single_rank(built):
    """world_size 2 over gloo: pack / exchange / unpack with the rank views equals the single-rank DSS bit for bit."""
    import torch.multiprocessing as mp
    from helpers import make_oracle
    ne, world, port = 4, 2, 29641
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dss_worker, args=(r, world, port, ne, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, v, hv, o = make_oracle(ne, 1, 11)
    rng = np.random.default_rng(5)
    field = rng.random((m.nelem, 3, 16))
    ref = np.ascontiguousarray(field * m.spheremp[:, None, :])
    o.dss(ref)
    ref *= m.rspheremp[:, None, :]
    got = np.zeros_like(ref)
    for rank, gid, out in parts:
        got[gid] = out
    assert np.array_equal(got, ref)


def test_abi_exports_every_declared_symbol():
    """The C-ABI library loads on a CPU-only box and exports every entry point include/tse.h declares (no compute calls)."""
    from transport_se_b200 import _build
    from transport_se_b200.advection import cuda_lib, EXPORTS
    _build.build_cuda()
    lib = cuda_lib()
    hdr = open(os.path.join(ROOT, "include", "tse.h")).read()
    declared = set(re.findall(r"\b(tse_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)
    lib.tse_last_error.restype = C.c_char_p
    assert lib.tse_device_count() >= 0


def test_no_cpu_fallback_without_device():
    """On a box without a GPU tse_init must fail loudly (there is no CPU path)."""
    from transport_se_b200.advection import cuda_lib, TracerAdvection, TseError
    from transport_se_b200.mesh import load_vcoord
    if cuda_lib().tse_device_count() > 0:
        pytest.skip("a CUDA device is present")
    m = mesh_for(4)
    with pytest.raises(TseError, match="no CUDA device"):
        TracerAdvection(m, m.local_view(), load_vcoord(), qsize=2)


def test_product_never_touches_the_oracle():
    """Nothing under transport_se_b200/ may import, link or load anything under oracle/."""
    pkg = os.path.join(ROOT, "transport_se_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if f == "_build.py":
                    continue  # build_oracle() compiles the checker; it does not use it
                assert "liboracle" not in txt and "oracle_lib" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)
