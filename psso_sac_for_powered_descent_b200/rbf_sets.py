"""Host-side tables for the *local* thin-plate-spline interpolants of C_D and C_L.

The reference evaluates `scipy.interpolate.RBFInterpolator(kernel='thin_plate_spline',
neighbors=50)` twice per physics sub-step (`src/envs/utils/aerodynamic_coefficients.py:
57-66`).  That interpolant is piecewise: for every query it takes the 50 nearest data
points, solves a 53x53 system for them and evaluates it - so the value jumps (up to
1.3e-2) wherever the 50-NN set changes, and a kernel that wants the reference's
numbers must use *exactly* scipy's neighbour set and scipy's coefficients.

Design (B200-first, not a translation of scipy):

* the data sit on a few AoA levels; on a level, squared distance to a query is convex
  in the Mach-sorted index, so any 50-NN set is one contiguous index interval
  [lo, hi) per level.  A set is therefore 2*L small integers (packed into a 64-bit key);
* every set that any query inside the reachable (Mach, AoA) box can produce is a cell
  of the order-50 Voronoi diagram.  `enumerate_sets` walks those cells exhaustively
  (breadth-first across cell edges, cells obtained by half-plane clipping), so the
  device never meets an unknown set;
* for each set the coefficients come from scipy's own build+solve routine (LAPACK
  dsysv on scipy's own matrix), bit-identical to what the reference computes on the
  fly, re-ordered into the kernel's level-major / Mach-ascending summation order;
* on the device a thread keeps its current set in registers, verifies it with 4
  distance comparisons per level and walks to the neighbouring cell when the query
  crosses an edge (csrc/pd_rbf.cuh); set id lookups go through an open-addressing hash
  table built here.
"""
from __future__ import annotations

import hashlib
import math
import os
from dataclasses import dataclass

import numpy as np

K_NEIGHBOURS = 50
COEF_STRIDE = 58       # 50 rbf + 3 poly + shift(2) + scale(2) = 57, padded to 58 doubles
MAX_LEVELS = 5


def _solve_sets(points, values, index_sets):
    """scipy's own coefficients for each neighbourhood (ascending original indices)."""
    from scipy.interpolate import RBFInterpolator
    probe = RBFInterpolator(points, values, kernel="thin_plate_spline", neighbors=K_NEIGHBOURS)
    try:
        from scipy.interpolate import _rbfinterp_np as backend
        solve = lambda idx: backend._build_and_solve_system(  # noqa: E731
            probe.y[idx], probe.d[idx], probe.smoothing[idx], probe.kernel, probe.epsilon,
            probe.powers, np)
    except Exception:  # older / newer scipy layouts: the public class does the same solve
        def solve(idx):
            it = RBFInterpolator(points[idx], values[idx], kernel="thin_plate_spline")
            return it._shift, it._scale, it._coeffs
    out = []
    for idx in index_sets:
        shift, scale, coeffs = solve(np.asarray(idx))
        out.append((np.asarray(shift, float), np.asarray(scale, float),
                    np.asarray(coeffs, float).reshape(-1)))
    return out


@dataclass
class LocalRbfTable:
    levels: np.ndarray        # [L] AoA value per level (ascending)
    level_off: np.ndarray     # [L+1] offsets into mach_sorted
    mach_sorted: np.ndarray   # [n] level-major, Mach ascending inside a level
    orig_index: np.ndarray    # [n] original data index of every sorted slot
    set_lo: np.ndarray        # [S, L] uint8
    set_hi: np.ndarray        # [S, L] uint8
    coeffs: np.ndarray        # [S, COEF_STRIDE] float64
    hash_keys: np.ndarray     # [H] uint64 (0 = empty)
    hash_vals: np.ndarray     # [H] int32
    box: tuple                # (mach_lo, mach_hi, aoa_lo, aoa_hi) the enumeration covered
    rows: np.ndarray = None   # [S, ROW_BYTES] uint8: device row = 57 doubles + 50 point indices
    points: np.ndarray = None  # [n, 2] (Mach, AoA) of every sorted slot
    grids: list = None        # QueryGrid per device lookup region

    @property
    def n_sets(self):
        return len(self.set_lo)

    # ------------------------------------------------------------------ host evaluation
    def find_set(self, mach, aoa):
        """Brute-force 50-NN of one query -> (lo[L], hi[L])."""
        L = len(self.levels)
        d2 = np.empty(len(self.mach_sorted))
        for l in range(L):
            a, b = self.level_off[l], self.level_off[l + 1]
            d2[a:b] = (self.mach_sorted[a:b] - mach) ** 2 + (self.levels[l] - aoa) ** 2
        sel = np.sort(np.argpartition(d2, K_NEIGHBOURS - 1)[:K_NEIGHBOURS])
        lo = np.zeros(L, int)
        hi = np.zeros(L, int)
        for l in range(L):
            a, b = self.level_off[l], self.level_off[l + 1]
            s = sel[(sel >= a) & (sel < b)] - a
            if len(s):
                lo[l], hi[l] = s[0], s[-1] + 1
                assert hi[l] - lo[l] == len(s), "non-contiguous neighbourhood"
            else:
                lo[l] = hi[l] = int(np.searchsorted(self.mach_sorted[a:b], mach))
        return lo, hi

    def lookup(self, lo, hi):
        key = pack_key(lo, hi)
        H = len(self.hash_keys)
        h = hash_u64(key) & (H - 1)
        while True:
            k = int(self.hash_keys[h])
            if k == key:
                return int(self.hash_vals[h])
            if k == 0:
                return -1
            h = (h + 1) & (H - 1)

    def evaluate(self, mach, aoa):
        """Host mirror of the device evaluation (used by the CPU-side self checks)."""
        lo, hi = self.find_set(mach, aoa)
        # empty levels carry no information in the key
        sid = self.lookup(lo, hi)
        if sid < 0:
            raise KeyError("neighbourhood not in table")
        c = self.coeffs[sid]
        acc = 0.0
        k = 0
        for l in range(len(self.levels)):
            a = self.level_off[l]
            for i in range(lo[l], hi[l]):
                r2 = (mach - self.mach_sorted[a + i]) ** 2 + (aoa - self.levels[l]) ** 2
                if r2 > 0.0:
                    acc += c[k] * (0.5 * r2 * np.log(r2))
                k += 1
        xh = (mach - c[53]) / c[55]
        yh = (aoa - c[54]) / c[56]
        return acc + c[50] + c[51] * xh + c[52] * yh


def pack_key(lo, hi):
    """6 bits per bound, empty levels normalised to lo=hi=0 so the key is canonical.
    Bit 63 is always set so that 0 can mean 'empty slot'."""
    key = 1 << 63
    for l in range(len(lo)):
        a, b = int(lo[l]), int(hi[l])
        if a == b:
            a = b = 0
        key |= (a & 63) << (12 * l)
        key |= (b & 63) << (12 * l + 6)
    return key


def hash_u64(k):
    """splitmix64 finaliser (same constants in csrc/pd_rbf.cuh)."""
    k &= (1 << 64) - 1
    k ^= k >> 30
    k = (k * 0xBF58476D1CE4E5B9) & ((1 << 64) - 1)
    k ^= k >> 27
    k = (k * 0x94D049BB133111EB) & ((1 << 64) - 1)
    k ^= k >> 31
    return k


# ---------------------------------------------------------------------------
def _clip(poly, labels, nx, ny, c, label):
    """Clip convex polygon (list of (x,y)) with half-plane nx*x+ny*y <= c.
    labels[i] tags the edge poly[i] -> poly[i+1]."""
    n = len(poly)
    if n == 0:
        return poly, labels
    d = [nx * p[0] + ny * p[1] - c for p in poly]
    if max(d) <= 0.0:
        return poly, labels
    if min(d) > 0.0:
        return [], []
    out, out_l = [], []
    for i in range(n):
        j = (i + 1) % n
        pi, pj, di, dj = poly[i], poly[j], d[i], d[j]
        if di <= 0.0:
            out.append(pi)
            if dj <= 0.0:
                out_l.append(labels[i])
            else:
                out_l.append(labels[i])
                t = di / (di - dj)
                out.append((pi[0] + t * (pj[0] - pi[0]), pi[1] + t * (pj[1] - pi[1])))
                out_l.append(label)
        elif dj <= 0.0:
            t = di / (di - dj)
            out.append((pi[0] + t * (pj[0] - pi[0]), pi[1] + t * (pj[1] - pi[1])))
            out_l.append(labels[i])
    return out, out_l


def _area(poly):
    s = 0.0
    for i in range(len(poly)):
        x0, y0 = poly[i]
        x1, y1 = poly[(i + 1) % len(poly)]
        s += x0 * y1 - x1 * y0
    return 0.5 * s


def enumerate_sets(mach_lv, levels, box, seeds):
    """All 50-NN sets reachable from queries inside `box`.

    mach_lv: list of ascending Mach arrays, one per level.  Returns a list of
    (lo, hi) tuples.  Exhaustive walk over the order-k Voronoi cells: a cell is the
    intersection of the half-planes 'in-set end point i is not farther than out-of-set
    frontier point j'; crossing edge (i, j) swaps i for j.
    """
    L = len(levels)
    nl = [len(m) for m in mach_lv]
    norm2 = [m * m + a * a for m, a in zip(mach_lv, levels)]
    m_lo, m_hi, a_lo, a_hi = box
    box_poly = [(m_lo, a_lo), (m_hi, a_lo), (m_hi, a_hi), (m_lo, a_hi)]

    def canon(lo, hi):
        lo = list(lo)
        hi = list(hi)
        for l in range(L):
            if lo[l] == hi[l]:
                lo[l] = hi[l] = 0
        return tuple(lo), tuple(hi)

    seen = {}
    queue = []
    for lo, hi in seeds:
        k = canon(lo, hi)
        if k not in seen:
            seen[k] = True
            queue.append(k)
    while queue:
        lo, hi = queue.pop()
        ins = []
        outs = []
        for l in range(L):
            if hi[l] > lo[l]:
                ins.append((l, lo[l]))
                if hi[l] - 1 != lo[l]:
                    ins.append((l, hi[l] - 1))
                if lo[l] > 0:
                    outs.append((l, lo[l] - 1))
                if hi[l] < nl[l]:
                    outs.append((l, hi[l]))
            else:
                outs.extend((l, j) for j in range(nl[l]))
        poly, labels = list(box_poly), [None] * 4
        for (li, i) in ins:
            mi, ai, ni = mach_lv[li][i], levels[li], norm2[li][i]
            for (lj, j) in outs:
                mj, aj, nj = mach_lv[lj][j], levels[lj], norm2[lj][j]
                poly, labels = _clip(poly, labels, 2 * (mj - mi), 2 * (aj - ai), nj - ni,
                                     ((li, i), (lj, j)))
                if not poly:
                    break
            if not poly:
                break
        if len(poly) < 3 or abs(_area(poly)) < 1e-13:
            seen[(lo, hi)] = False
            continue
        for lab in set(l for l in labels if l is not None):
            (li, i), (lj, j) = lab
            nlo, nhi = list(lo), list(hi)
            if i == lo[li]:
                nlo[li] += 1
            else:
                nhi[li] -= 1
            if nhi[lj] > nlo[lj]:
                if j == nlo[lj] - 1:
                    nlo[lj] -= 1
                elif j == nhi[lj]:
                    nhi[lj] += 1
                else:       # swap emptied this level first (li == lj cannot do that)
                    continue
            else:
                nlo[lj], nhi[lj] = j, j + 1
            k = canon(nlo, nhi)
            if k not in seen:
                seen[k] = True
                queue.append(k)
    return [k for k, ok in seen.items() if ok]


ROW_BYTES = 512        # 57 doubles (coeffs, poly, shift, scale) + 50 point-index bytes, padded


@dataclass
class QueryGrid:
    """Uniform (Mach, AoA) grid over one query box.  cells[ia*nm+im] >= 0: the whole grid cell
    lies inside one order-50 Voronoi cell (all four corners and the centre give the same set;
    the Voronoi cells are convex) and the value is that set's id - no search needed on the
    device.  cells < 0: -(k+1) indexes (imp_hint[k], imp_id[k]), the set at the cell centre,
    from which the device walks to the exact set."""
    m0: float
    dm: float
    nm: int
    a0: float
    da: float
    na: int
    cells: np.ndarray
    imp_hint: np.ndarray
    imp_id: np.ndarray

    @property
    def pure_fraction(self):
        return float(np.mean(self.cells >= 0))


def _sets_of_queries(tbl, M, A):
    """Vectorised brute-force 50-NN for many queries -> (lo[n,L], hi[n,L]) with empty levels
    parked at the insertion point of M."""
    L = len(tbl.levels)
    n = len(M)
    lo = np.zeros((n, L), np.int64)
    hi = np.zeros((n, L), np.int64)
    lvl_of = np.repeat(np.arange(L), np.diff(tbl.level_off))
    a_pts = tbl.levels[lvl_of]
    for s in range(0, n, 8192):
        m, a = M[s:s + 8192, None], A[s:s + 8192, None]
        d2 = (tbl.mach_sorted[None, :] - m) ** 2 + (a_pts[None, :] - a) ** 2
        sel = np.argpartition(d2, K_NEIGHBOURS - 1, axis=1)[:, :K_NEIGHBOURS]
        mask = np.zeros(d2.shape, bool)
        np.put_along_axis(mask, sel, True, axis=1)
        for l in range(L):
            o0, o1 = tbl.level_off[l], tbl.level_off[l + 1]
            ml = mask[:, o0:o1]
            cnt = ml.sum(1)
            first = np.argmax(ml, axis=1)
            ins = np.searchsorted(tbl.mach_sorted[o0:o1], M[s:s + 8192])
            lo[s:s + 8192, l] = np.where(cnt > 0, first, ins)
            hi[s:s + 8192, l] = np.where(cnt > 0, first + cnt, ins)
    return lo, hi


def pack_keys(lo, hi):
    """Vectorised pack_key for [n, L] bound arrays -> uint64[n]."""
    lo = np.asarray(lo, np.uint64)
    hi = np.asarray(hi, np.uint64)
    empty = lo == hi
    a = np.where(empty, np.uint64(0), lo & np.uint64(63))
    b = np.where(empty, np.uint64(0), hi & np.uint64(63))
    key = np.full(lo.shape[0], np.uint64(1) << np.uint64(63), np.uint64)
    for l in range(lo.shape[1]):
        key |= (a[:, l] << np.uint64(12 * l)) | (b[:, l] << np.uint64(12 * l + 6))
    return key


def pack_hints(lo, hi):
    """Vectorised pack_hint for [n, L] bound arrays -> uint64[n]."""
    lo = np.asarray(lo, np.uint64) & np.uint64(63)
    hi = np.asarray(hi, np.uint64) & np.uint64(63)
    h = np.zeros(lo.shape[0], np.uint64)
    for l in range(lo.shape[1]):
        h |= (lo[:, l] << np.uint64(12 * l)) | (hi[:, l] << np.uint64(12 * l + 6))
    return h


def pack_hint(lo, hi):
    h = 0
    for l in range(len(lo)):
        h |= (int(lo[l]) & 63) << (12 * l)
        h |= (int(hi[l]) & 63) << (12 * l + 6)
    return h


def build_grid(tbl, box, nm, na) -> QueryGrid:
    m0, m1, a0, a1 = box
    dm = (m1 - m0) / nm
    da = (a1 - a0) / na if na > 1 else max(a1 - a0, 1e-9)
    set_keys = pack_keys(tbl.set_lo, tbl.set_hi)
    order = np.argsort(set_keys)
    sorted_keys = set_keys[order]

    def ids(M, A):
        lo, hi = _sets_of_queries(tbl, M, A)
        k = pack_keys(lo, hi)
        pos = np.minimum(np.searchsorted(sorted_keys, k), len(sorted_keys) - 1)
        out = np.where(sorted_keys[pos] == k, order[pos], -1).astype(np.int64)
        return out, lo, hi
    # corners (na+1) x (nm+1)
    cm = m0 + dm * np.arange(nm + 1)
    ca = a0 + da * np.arange(na + 1) if na > 1 else np.array([a0, a1])
    CM, CA = np.meshgrid(cm, ca)
    cid, _, _ = ids(CM.ravel(), CA.ravel())
    cid = cid.reshape(len(ca), nm + 1)
    mm = m0 + dm * (np.arange(nm) + 0.5)
    aa = a0 + da * (np.arange(na) + 0.5) if na > 1 else np.array([0.5 * (a0 + a1)])
    MM, AA = np.meshgrid(mm, aa)
    zid, zlo, zhi = ids(MM.ravel(), AA.ravel())
    assert (cid >= 0).all() and (zid >= 0).all(), "query grid meets a set missing from the table"
    zid2 = zid.reshape(len(aa), nm)
    pure = (cid[:-1, :-1] == zid2) & (cid[:-1, 1:] == zid2) & (cid[1:, :-1] == zid2) & (cid[1:, 1:] == zid2)
    cells = np.where(pure, zid2, 0).astype(np.int32).ravel()
    imp = np.nonzero(~pure.ravel())[0]
    cells[imp] = -(np.arange(len(imp), dtype=np.int32) + 1)
    imp_hint = pack_hints(zlo[imp], zhi[imp]) if len(imp) else np.zeros(0, np.uint64)
    imp_id = zid[imp].astype(np.int32)
    if len(imp):
        _resolve_single_edge_cells(tbl, ids, imp, cid, zid2, m0, dm, a0, da, nm, imp_hint, imp_id)
    return QueryGrid(float(m0), float(dm), int(nm), float(a0), float(da), int(len(aa)), cells,
                     imp_hint, imp_id)


EDGE_FLAG = np.uint64(1) << np.uint64(63)


def _resolve_single_edge_cells(tbl, ids, imp, cid, zid2, m0, dm, a0, da, nm, imp_hint, imp_id):
    """Impure grid cells that are cut by exactly ONE order-50 Voronoi edge need no search on the
    device: the two sets differ by one swap (S2 = S1 - {p} + {q}), the edge is the bisector of p
    and q, and the query belongs to S1 iff it is not farther from p than from q.  Such a cell is
    re-encoded in place: imp_id = S1, imp_hint = bit 63 | S2 << 16 | p << 8 | q (p, q = point
    slots).  Proof obligation per cell (all checked here with the brute-force 50-NN): the five
    samples (corners + centre) show only S1 and S2, each on its side of the bisector, and the two
    points where the bisector crosses the cell boundary, nudged to either side, show S1 / S2 -
    every vertex of the two convex polygons the bisector cuts the cell into lies in the (convex)
    Voronoi cell of its set, hence so do the polygons."""
    L = len(tbl.levels)
    S, NP = tbl.n_sets, len(tbl.mach_sorted)
    member = np.zeros((S, NP), bool)
    lo_all, hi_all = np.asarray(tbl.set_lo, np.int64), np.asarray(tbl.set_hi, np.int64)
    for l in range(L):
        idx = np.arange(tbl.level_off[l + 1] - tbl.level_off[l])[None, :]
        member[:, tbl.level_off[l]:tbl.level_off[l + 1]] = (idx >= lo_all[:, l:l + 1]) & (idx < hi_all[:, l:l + 1])
    ia, im = np.divmod(imp, nm)
    five = np.stack([cid[ia, im], cid[ia, im + 1], cid[ia + 1, im], cid[ia + 1, im + 1], zid2[ia, im]], 1)
    s1, s2 = five.min(1), five.max(1)
    cand = ((five == s1[:, None]) | (five == s2[:, None])).all(1) & (s1 != s2)
    k = np.nonzero(cand)[0]
    if len(k) == 0:
        return
    d1 = member[s1[k]] & ~member[s2[k]]
    d2 = member[s2[k]] & ~member[s1[k]]
    single = (d1.sum(1) == 1) & (d2.sum(1) == 1)
    k, d1, d2 = k[single], d1[single], d2[single]
    p, q = d1.argmax(1), d2.argmax(1)
    pts = np.asarray(tbl.points, float).reshape(-1, 2)
    P, Q = pts[p], pts[q]
    nrm = Q - P
    c = 0.5 * ((Q ** 2).sum(1) - (P ** 2).sum(1))
    x0, y0 = m0 + dm * im[k], a0 + da * ia[k]
    corners = np.stack([np.stack([x0, y0], 1), np.stack([x0 + dm, y0], 1), np.stack([x0, y0 + da], 1),
                        np.stack([x0 + dm, y0 + da], 1), np.stack([x0 + 0.5 * dm, y0 + 0.5 * da], 1)], 1)
    side = (corners * nrm[:, None, :]).sum(2) - c[:, None]            # < 0: nearer p -> S1
    want = np.where(side < 0, s1[k][:, None], s2[k][:, None])
    ok = (side != 0).all(1) & (want == five[k]).all(1)
    # the two boundary crossings: rectangle edges in order 00-01, 01-11, 11-10, 10-00
    order = [(0, 1), (1, 3), (3, 2), (2, 0)]
    cross = np.stack([(side[:, a] < 0) != (side[:, b] < 0) for a, b in order], 1)
    ok &= cross.sum(1) == 2
    k, p, q, nrm, corners, side, cross = k[ok], p[ok], q[ok], nrm[ok], corners[ok], side[ok], cross[ok]
    if len(k) == 0:
        return
    X = []
    for e, (a, b) in enumerate(order):
        t = side[:, a] / (side[:, a] - side[:, b] + (side[:, a] == side[:, b]))
        X.append(corners[:, a] + (corners[:, b] - corners[:, a]) * t[:, None])
    X = np.stack(X, 1)                                                 # [n, 4 edges, 2]
    first = cross.argmax(1)
    second = 3 - cross[:, ::-1].argmax(1)
    n_hat = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    eps = 1e-6 * math.hypot(dm, da)
    good = np.ones(len(k), bool)
    for sel in (first, second):
        x = X[np.arange(len(k)), sel]
        for sgn, sid in ((-1.0, s1[k]), (1.0, s2[k])):
            y = x + sgn * eps * n_hat
            got, _, _ = ids(y[:, 0], y[:, 1])
            good &= got == sid
    k, p, q = k[good], p[good], q[good]
    imp_id[k] = s1[k].astype(np.int32)
    imp_hint[k] = EDGE_FLAG | (s2[k].astype(np.uint64) << np.uint64(16)) | (p.astype(np.uint64) << np.uint64(8)) \
        | q.astype(np.uint64)


def build_table(mach, aoa, val, boxes, cache_dir=None, grid=(64, 24), grids=None) -> LocalRbfTable:
    """`boxes`: list of (mach_lo, mach_hi, aoa_lo, aoa_hi) query regions whose neighbour sets
    are enumerated.  `grids`: list of (box, nm, na) device lookup grids."""
    import pickle
    mach = np.asarray(mach, float)
    aoa = np.asarray(aoa, float)
    val = np.asarray(val, float)
    box = np.asarray(boxes, float).reshape(-1)
    tag = hashlib.sha256(mach.tobytes() + aoa.tobytes() + val.tobytes() + box.tobytes()
                         + repr(grids).encode() + b"v10").hexdigest()[:16]
    if cache_dir:
        path = os.path.join(cache_dir, f"rbf_{tag}.pkl")
        if os.path.exists(path):
            with open(path, "rb") as f:
                return pickle.load(f)
    levels = np.unique(aoa)
    assert len(levels) <= MAX_LEVELS
    mach_lv, orig_lv = [], []
    for a in levels:
        idx = np.nonzero(aoa == a)[0]
        order = np.argsort(mach[idx], kind="stable")
        mach_lv.append(mach[idx][order])
        orig_lv.append(idx[order])
        assert len(idx) < 64
    level_off = np.concatenate([[0], np.cumsum([len(m) for m in mach_lv])]).astype(np.int32)
    mach_sorted = np.concatenate(mach_lv)
    orig_index = np.concatenate(orig_lv).astype(np.int32)
    assert len(mach_sorted) < 256
    tbl = LocalRbfTable(levels, level_off, mach_sorted, orig_index, None, None, None, None, None,
                        tuple(box))
    # seeds: a coarse brute-force grid per box, then the exhaustive cell walk
    found = {}
    for bx in boxes:
        seeds = {}
        for a in np.linspace(bx[2], bx[3], grid[1]):
            for m in np.linspace(bx[0], bx[1], grid[0]):
                lo, hi = tbl.find_set(m, a)
                seeds[(tuple(lo), tuple(hi))] = True
        for k in enumerate_sets(mach_lv, list(levels), bx, list(seeds.keys())):
            found[k] = True
    sets = sorted(found.keys())
    S = len(sets)
    set_lo = np.array([s[0] for s in sets], np.uint8).reshape(S, len(levels))
    set_hi = np.array([s[1] for s in sets], np.uint8).reshape(S, len(levels))
    # scipy coefficients per set, permuted into level-major / Mach-ascending order
    points = np.column_stack((mach, aoa))
    index_sets, perms, mine_all = [], [], []
    for lo, hi in sets:
        mine = np.concatenate([orig_lv[l][lo[l]:hi[l]] for l in range(len(levels))])
        assert len(mine) == K_NEIGHBOURS
        asc = np.sort(mine)
        index_sets.append(asc)
        perms.append(np.searchsorted(asc, mine))     # position of my k-th point in scipy's order
        mine_all.append(np.concatenate([level_off[l] + np.arange(lo[l], hi[l]) for l in range(len(levels))]))
    solved = _solve_sets(points, val, index_sets)
    coeffs = np.zeros((S, COEF_STRIDE))
    rows = np.zeros((S, ROW_BYTES), np.uint8)
    for s, ((shift, scale, c), perm) in enumerate(zip(solved, perms)):
        coeffs[s, :50] = c[:50][perm]
        coeffs[s, 50:53] = c[50:53]
        coeffs[s, 53:55] = shift
        coeffs[s, 55:57] = scale
        dev = coeffs[s, :57].copy()
        dev[:50] *= 0.5            # r^2 log r = (c/2) r^2 log r^2 ; exact scaling
        dev[55:57] = 1.0 / dev[55:57]     # device multiplies by 1/scale
        rows[s, :57 * 8] = dev.view(np.uint8)
        rows[s, 57 * 8:57 * 8 + 50] = mine_all[s].astype(np.uint8)
    H = 1
    while H < 4 * S:
        H *= 2
    hk = np.zeros(H, np.uint64)
    hv = np.full(H, -1, np.int32)
    for s, (lo, hi) in enumerate(sets):
        key = pack_key(lo, hi)
        h = hash_u64(key) & (H - 1)
        while hk[h] != 0:
            h = (h + 1) & (H - 1)
        hk[h] = key
        hv[h] = s
    tbl.set_lo, tbl.set_hi, tbl.coeffs, tbl.hash_keys, tbl.hash_vals = set_lo, set_hi, coeffs, hk, hv
    tbl.rows = rows
    lvl_of = np.repeat(np.arange(len(levels)), np.diff(level_off))
    tbl.points = np.ascontiguousarray(np.column_stack((mach_sorted, levels[lvl_of])))
    tbl.grids = [build_grid(tbl, bx, nm, na) for (bx, nm, na) in (grids or [])]
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        with open(path, "wb") as f:
            pickle.dump(tbl, f)
    return tbl


# Reachable query boxes: Mach is clamped to [0, 10]
# (rockets_physics.py:481).  C_D is queried at clamp(degrees(alpha_eff), +-radians(10))
# (aerodynamic_coefficients.py:108-114); C_L at |aoa| in (1e-6, 10] plus the
# un-negated (Mach, -10) branch (aerodynamic_coefficients.py:120-131).
CD_BOXES = [(-0.02, 10.02, -0.2, 0.2)]
CL_BOXES = [(-0.02, 10.02, -0.01, 10.01), (-0.02, 10.02, -10.01, -9.99)]
_LIM = 0.17453292519943295 * (1 + 1e-9)
# device lookup grids: (box, n_mach, n_aoa).  Mach is clamped to [0, 10] upstream.  The grids must
# cover that whole range: beyond the last data point (Mach 5.51) the 50-NN set still changes
# (cross-level distance orderings keep flipping as Mach grows), so no column can stand in for the
# tail.  C_L 2-D grid: 4096 x 512 cells (8 MB, L2-resident) are 83 % pure against 73 % for
# 1024 x 256, +3 % step throughput on B200 (fewer lanes take the exact walk).
CD_GRIDS = [((0.0, 10.0, -_LIM, _LIM), 2048, 8)]
CL_GRIDS = [((0.0, 10.0, 0.0, 10.0), 4096, 512), ((0.0, 10.0, -10.0 - 1e-9, -10.0 + 1e-9), 2048, 1)]
if os.environ.get("PD_CL_GRID"):      # experiment knob: "mach_max,n_mach,n_aoa"
    _mm, _nm, _na = os.environ["PD_CL_GRID"].split(",")
    CL_GRIDS[0] = ((0.0, float(_mm), 0.0, 10.0), int(_nm), int(_na))
