"""Geometry unit tests of the box-box routine: the kernel source (fp32, host emulation) against the checker (fp64) on
random box pairs, and both against first principles (every reported point lies on / inside both boxes up to the
reported penetration, the normal is a unit vector pointing from box 1 to box 2, separated boxes give no contact)."""
import ctypes as C

import numpy as np

from host_emulation import emu
from oracle import engine


def _rot(rs):
    q = rs.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _inside(p, c, R, s, tol):
    loc = R.T @ (p - c)
    return np.all(np.abs(loc) <= s + tol)


def test_box_box_kernel_vs_checker_and_first_principles():
    Lo, Le = engine.lib(), emu.lib()
    D = np.ctypeslib.ndpointer(np.float64, flags="C")
    F = np.ctypeslib.ndpointer(np.float32, flags="C")
    Lo.mje_test_box_box.argtypes = [D, D, D, D, D, D, C.c_double, D]
    Le.emu_box_box.argtypes = [F, F, F, F, F, F, C.c_float, F]
    rs = np.random.RandomState(0)
    hits = agree = 0
    for trial in range(3000):
        s1, s2 = rs.uniform(0.01, 0.2, 3), rs.uniform(0.01, 0.2, 3)
        R1, R2 = _rot(rs), _rot(rs)
        p1 = np.zeros(3)
        p2 = rs.normal(size=3)
        p2 *= rs.uniform(0.02, 0.35) / np.linalg.norm(p2)
        margin = 0.001 if trial % 2 else 0.0
        o64 = np.zeros((8, 7))
        n64 = Lo.mje_test_box_box(p1, R1.ravel().copy(), s1, p2, R2.ravel().copy(), s2, margin, o64)
        f = lambda a: np.ascontiguousarray(a, np.float32)  # noqa: E731
        o32 = np.zeros((8, 7), np.float32)
        n32 = Le.emu_box_box(f(p1), f(R1.ravel()), f(s1), f(p2), f(R2.ravel()), f(s2), margin, o32)
        if n64 == 0:
            # separated (beyond the margin): no corner of either box may be inside the other
            for (c, R, s, oc, oR, os_) in ((p1, R1, s1, p2, R2, s2), (p2, R2, s2, p1, R1, s1)):
                for v in range(8):
                    corner = c + R @ (s * [(v & 1) * 2 - 1, (v >> 1 & 1) * 2 - 1, (v >> 2 & 1) * 2 - 1])
                    assert not _inside(corner, oc, oR, os_, -1e-9)
            assert n32 == 0 or abs(o32[:n32, 6].max() - margin) < 1e-5
            continue
        hits += 1
        for c in range(n64):
            pos, nrm, dist = o64[c, :3], o64[c, 3:6], o64[c, 6]
            assert abs(np.linalg.norm(nrm) - 1) < 1e-9 and nrm @ (p2 - p1) > -1e-9 and dist < margin
            # the contact point sits halfway between the two surfaces: within half the penetration (+ margin) of both boxes;
            # a face contact reports dist = -depth / 2 (MuJoCo 2.1.0, pinned by the peg landing), an edge contact -depth
            tol = abs(dist) + margin + 1e-7
            assert _inside(pos, p1, R1, s1, tol) and _inside(pos, p2, R2, s2, tol)
        if n32 == n64:
            d32, d64 = np.sort(o32[:n32, 6]), np.sort(o64[:n64, 6])
            if np.abs(d32 - d64).max() < 2e-5 and np.abs(np.abs(o32[0, 3:6] @ o64[0, 3:6]) - 1) < 1e-4:
                agree += 1
    assert hits > 500
    # fp32 and fp64 pick the same axis and the same clipped polygon except at near-ties between candidate axes
    assert agree >= 0.97 * hits, (agree, hits)


def _support(tp, c, R, s, d):
    """Support point of a box (type 6) or a z-axis cylinder (type 5, s = (radius, half height)) in direction d."""
    dl = R.T @ d
    if tp == 6:
        loc = np.sign(dl) * s
        loc[dl == 0] = s[dl == 0]
    else:
        t = np.hypot(dl[0], dl[1])
        loc = np.array([dl[0] / t * s[0], dl[1] / t * s[0], np.sign(dl[2]) * s[1]]) if t > 1e-15 else np.array([0, 0, np.sign(dl[2]) * s[1]])
    return c + R @ loc


def test_mpr_kernel_vs_checker_and_first_principles():
    """Minkowski portal refinement on random box / cylinder pairs: the reported (direction, depth) must be a genuine
    overlap of the two shapes along that direction (support-function identity, to the portal tolerance), separated
    shapes must give no contact, and the fp32 kernel routine must agree with the fp64 checker wherever the latter's
    answer is well conditioned."""
    Lo, Le = engine.lib(), emu.lib()
    D = np.ctypeslib.ndpointer(np.float64, flags="C")
    F = np.ctypeslib.ndpointer(np.float32, flags="C")
    Lo.mje_test_mpr.argtypes = [C.c_int, D, D, D, C.c_int, D, D, D, C.c_double, D]
    Le.emu_mpr.argtypes = [C.c_int, F, F, F, C.c_int, F, F, F, C.c_float, F]
    rs = np.random.RandomState(1)
    hits = agree = 0
    for trial in range(2500):
        t1, t2 = (5, 6) if trial % 3 else (6, 6)      # MuJoCo orders a pair by geom type: cylinder (5) before box (6)
        s1 = np.array([rs.uniform(0.01, 0.05), rs.uniform(0.02, 0.1), 0.0]) if t1 == 5 else rs.uniform(0.01, 0.1, 3)
        s2 = rs.uniform(0.01, 0.1, 3)
        R1, R2 = _rot(rs), _rot(rs)
        p1 = np.zeros(3)
        p2 = rs.normal(size=3)
        p2 *= rs.uniform(0.01, 0.2) / np.linalg.norm(p2)
        o64 = np.zeros(7)
        n64 = Lo.mje_test_mpr(t1, s1, p1, R1.ravel().copy(), t2, s2, p2, R2.ravel().copy(), 0.0, o64)
        f = lambda a: np.ascontiguousarray(a, np.float32)  # noqa: E731
        o32 = np.zeros(7, np.float32)
        n32 = Le.emu_mpr(t1, f(s1), f(p1), f(R1.ravel()), t2, f(s2), f(p2), f(R2.ravel()), 0.0, o32)
        if not n64:
            continue
        hits += 1
        depth, d = o64[0], o64[1:4]
        assert abs(np.linalg.norm(d) - 1) < 1e-9 and depth > 0
        # overlap of the shapes along d: h_1(d) - min over shape 2 of x.d
        overlap = _support(t1, p1, R1, s1, d) @ d - _support(t2, p2, R2, s2, -d) @ d
        assert overlap > 0 and depth <= overlap + 1e-5, (depth, overlap)
        if n32 and abs(o32[0] - depth) < 1e-4 and o32[1:4] @ d > 1 - 1e-4:
            agree += 1
    assert hits > 400
    assert agree >= 0.9 * hits, (agree, hits)
