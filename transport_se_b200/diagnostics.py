"""Verification tooling on the host: unique-column extraction and the DCMIP error norms.

Restates, in numpy, what the reference does outside the hot path to judge a run:
  * unique-point ownership -- a shared GLL node belongs to the element with the lowest
    global id (reference src/share/dof_mod.F90:42-59,322-357);
  * the L1/L2/Linf, q_max, q_min formulas of
    test/dcmip1-1/dcmip1-1_error_norm_ng.ncl:41-77 and test/dcmip1-2/dcmip1-2_error_norm_ng.ncl:42-77
    (including NCL's single-precision pi and the dz rebuilt from mid-level heights).
"""
import numpy as np

_EDGE_NODES = {2: [0, 1, 2, 3], 1: [3, 7, 11, 15], 3: [12, 13, 14, 15], 0: [0, 4, 8, 12]}  # SOUTH,EAST,NORTH,WEST
_CORNER_NODE = {4: 0, 5: 3, 6: 12, 7: 15}  # SWEST,SEAST,NWEST,NEAST


def unique_mask(mesh):
    """bool[nelem,16]: True where element e owns node n (lowest global id among sharers)."""
    n = mesh.nelem
    own = np.ones((n, 16), bool)
    ids = np.arange(n)
    for d, nodes in _EDGE_NODES.items():
        lose = mesh.nbr[:, d] < ids
        for nd in nodes:
            own[lose, nd] = False
    for d, nd in _CORNER_NODE.items():
        b = mesh.nbr[:, d]
        lose = (b >= 0) & (b < ids)
        own[lose, nd] = False
    return own


def dcmip_error_norms(mesh, q_i, q_f, z_mid, mask=None):
    """q_i, q_f: [nelem, nlev, 16] mixing ratio at t=0 and at the end; z_mid: [nlev] mid-level heights (geo/g).

    Returns dict(L1, L2, Linf, q_max, q_min) as the NCL scripts print them."""
    if mask is None:
        mask = unique_mask(mesh)
    pi32 = np.arccos(np.float32(-1.0))
    rad = np.float32(pi32 / np.float32(180.0))
    lat_deg = mesh.lat * (180.0 / np.pi)  # history file stores degrees
    lat = (lat_deg * np.float64(rad))[mask]  # [ncol]
    qi = np.transpose(q_i, (1, 0, 2))[:, mask]  # [nlev, ncol]
    qf = np.transpose(q_f, (1, 0, 2))[:, mask]
    nlev = qi.shape[0]
    dh = np.zeros(nlev)
    base = 0.0
    for i in range(1, nlev + 1):
        dh[nlev - i] = 2.0 * (z_mid[nlev - i] - base)
        base = base + dh[nlev - i]
    R = np.float64(np.float32(6.37122e6))
    dlat = 0.5 * np.float64(pi32) / (mesh.ne * 3)
    dV = (R * np.cos(lat) * dlat)[None, :] * (R * dlat) * dh[:, None]
    dq = qf - qi
    dev = np.abs(qi - qi.mean())
    return dict(L1=float(np.sum(np.abs(dq) * dV) / np.sum(dev * dV)),
                L2=float(np.sqrt(np.sum(dq ** 2 * dV)) / np.sqrt(np.sum(dev ** 2 * dV))),
                Linf=float(np.max(np.abs(dq) * dV) / np.max(dev * dV)),
                q_max=float(qf.max()), q_min=float(qf.min()))
