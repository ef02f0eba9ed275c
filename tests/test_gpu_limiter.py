"""The CUDA limiter (limiter_y in csrc/tse_tile.cuh: y-space, pre-check fast path, direction-fixed fused sweeps with carried
weights) against the oracle's line-by-line restatement of limiter_optim_iter_full (reference
src/share/prim_advection_mod.F90:976-1094) on adversarial planes, through the C ABI (tse_debug_limiter).

Tolerance: 1e-12 relative max-norm per plane (the north-star bar for one step); the bounds (in/out) to 1e-14."""
import numpy as np
import pytest



def _oracle(pt, sph, dpm, mn, mx):
    from oracle.oracle_lib import lib, _p
    L = lib()
    out = pt.copy()
    omn, omx = mn.copy(), mx.copy()
    for i in range(pt.shape[0]):
        p = np.ascontiguousarray(out[i])
        a, b = omn[i:i + 1].copy(), omx[i:i + 1].copy()
        L.orc_limiter_optim_iter_full(_p(p), _p(np.ascontiguousarray(sph[i])), _p(a), _p(b), _p(np.ascontiguousarray(dpm[i])))
        out[i], omn[i], omx[i] = p, a[0], b[0]
    return out, omn, omx


def _cases():
    rng = np.random.default_rng(20261018)
    pt, sph, dpm, mn, mx, tag = [], [], [], [], [], []

    def add(q, lo, hi, name, s=None, d=None):
        d = 1.0 + rng.random(16) if d is None else d
        s = (0.5 + rng.random(16)) * 1e-3 if s is None else s
        pt.append(np.asarray(q, dtype=np.float64) * d); sph.append(s); dpm.append(d); mn.append(lo); mx.append(hi); tag.append(name)

    gll = np.outer([1, 5, 5, 1], [1, 5, 5, 1]).ravel() / 36.0
    for t in range(300):  # random planes, bounds cut through the data: 1..several sweeps, both directions
        q = rng.random(16) * (3.0 if t % 3 else 0.3)
        add(q, 0.2, 0.8, "random")
    for t in range(60):   # infeasible bounds: mass/sumc outside [minp, maxp] -> relaxation (:1024-1029)
        add(rng.random(16), 0.9, 1.0, "infeasible-high")
        add(0.5 + rng.random(16), 0.0, 0.3, "infeasible-low")
    for t in range(40):   # min == max
        add(rng.random(16), 0.5, 0.5, "min==max")
    for t in range(40):   # every node outside the bounds on the same side / on both sides
        add(2.0 + rng.random(16), 0.0, 1.0, "all-above")
        add(-1.0 - rng.random(16), 0.0, 1.0, "all-below")
        add(np.where(rng.random(16) < 0.5, 2.0, -1.0) + 0.1 * rng.random(16), 0.0, 1.0, "all-clipped")
    for t in range(60):   # up to np*np-1 = 15 sweeps: one node far above the bound, the others stacked below it so that every
        w = gll * (1 + 0.1 * rng.random(16))     # redistribution pushes exactly one more node over (gaps from the recursion)
        d = 1.0 + rng.random(16)
        c = w * d
        order = rng.permutation(16)
        big, rest = order[0], order[1:1 + (t % 15) + 1]   # 1..15 receivers -> 2..15 sweeps
        gap = np.zeros(16)
        remaining = 50.0 * c.sum()                         # excess mass on the big node
        excess = remaining / c[big]
        cum, theta = 0.0, 0.01
        others = [n for n in order[1:] if n not in rest]
        for i, n in enumerate(rest):
            W = c[rest[i:]].sum()
            inc = remaining / W
            last = i == len(rest) - 1
            gap[n] = cum + (2.0 if last else theta) * inc   # the last receiver has room for all that is left: bounds stay feasible
            remaining = 0.0 if last else (inc - theta * inc) * c[n]
            cum += inc
        q = 1.0 - gap
        q[big] = 1.0 + excess
        q[others] = 1.0                                    # already at the bound
        add(q, -1e9, 1.0, "many-sweeps-up", s=w, d=d)
        add(2.0 - q, 1.0, 1e9, "many-sweeps-down", s=w, d=d)   # mirror image about 1: mass stays O(sum c), no cancellation
    for t in range(20):   # sumc <= 0 (:1016): plane returned untouched
        add(rng.random(16), 0.2, 0.8, "sumc<=0", d=-(1.0 + rng.random(16)))
    for t in range(20):   # constant field sitting on the bounds up to roundoff (interior of a checkerboard cell)
        add(1.0 + 1e-16 * rng.integers(-3, 4, 16), 1.0, 1.0, "roundoff-on-bound")
        add(np.zeros(16), 0.0, 0.0, "zero")
    for t in range(20):   # corner-weighted: only low-weight nodes can take the mass
        q = np.full(16, 1.0); q[[0, 3, 12, 15]] = 0.2 * rng.random(4); q[5] = 1.5 + t
        add(q, 0.0, 1.0, "corner-receivers", s=gll.copy())
    return (np.array(pt), np.array(sph), np.array(dpm), np.array(mn, dtype=np.float64), np.array(mx, dtype=np.float64), tag)


@pytest.mark.gpu
def test_limiter_matches_oracle_on_adversarial_planes(built):
    from transport_se_b200.advection import debug_limiter
    pt, sph, dpm, mn, mx, tag = _cases()
    ref, rmn, rmx = _oracle(pt, sph, dpm, mn, mx)
    got, gmn, gmx = debug_limiter(pt, sph, dpm, mn, mx)
    want = sph * ref
    scale = np.maximum(np.max(np.abs(want), axis=1), 1e-300)
    err = np.max(np.abs(got - want), axis=1) / scale
    worst = int(np.argmax(err))
    assert err[worst] < 1e-12, (tag[worst], err[worst])
    # planes the reference leaves untouched (sumc <= 0) come back bit for bit
    neg = np.array([t == "sumc<=0" for t in tag])
    assert np.array_equal(got[neg], (sph * pt)[neg])
    bscale = np.maximum(np.abs(rmx), np.abs(rmn)) + 1e-300
    assert np.max(np.abs(gmn - rmn) / bscale) < 1e-14 and np.max(np.abs(gmx - rmx) / bscale) < 1e-14
    # invariants of the device result itself: mass conserved to the limiter's tolerance, bounds respected
    c = sph * dpm
    ok = c.sum(axis=1) > 0
    mass0 = np.sum(sph * pt, axis=1)
    mass1 = np.sum(got, axis=1)
    assert np.max(np.abs(mass1 - mass0)[ok] / np.maximum(np.abs(mass0[ok]), 1e-300)) < 2e-13
    xg = got / c
    inb = ok
    assert np.all(xg[inb].min(axis=1) >= gmn[inb] - 1e-13 * np.maximum(1, np.abs(gmn[inb])))
    assert np.all(xg[inb].max(axis=1) <= gmx[inb] + 1e-13 * np.maximum(1, np.abs(gmx[inb])))


def test_limiter_sweep_counts_cover_1_to_15(built):
    """The many-sweep family must actually reach the reference's iteration cap (np*np - 1 = 15): count the oracle's sweeps
    by re-running its algorithm in numpy."""
    pt, sph, dpm, mn, mx, tag = _cases()
    counts = []
    for i in range(pt.shape[0]):
        c = sph[i] * dpm[i]
        if c.sum() <= 0:
            continue
        x = pt[i] / dpm[i]
        mass = np.sum(c * x)
        lo, hi = mn[i], mx[i]
        if mass < lo * c.sum(): lo = mass / c.sum()
        if mass > hi * c.sum(): hi = mass / c.sum()
        it = 0
        for it in range(1, 16):
            add = np.sum(np.where(x > hi, (x - hi) * c, 0)) - np.sum(np.where(x < lo, (lo - x) * c, 0))
            x = np.clip(x, lo, hi)
            if abs(add) <= 5e-14 * abs(mass):
                break
            if add > 0:
                m = x < hi
            else:
                m = x > lo
            if c[m].sum() == 0:
                break
            x = np.where(m, x + add / c[m].sum(), x)
        counts.append(it)
    counts = np.array(counts)
    assert counts.max() == 15 and (counts == 1).any() and len(np.unique(counts)) >= 8, np.bincount(counts)
