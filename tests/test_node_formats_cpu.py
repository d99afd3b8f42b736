"""The conservative box tests of the traversal kernels, restated in numpy float32 and checked against exact arithmetic.

`vkrt_trace.cu` decides hits with the pinned triangle test only; the box tests of all node formats merely have to be
*conservative* (never reject a box the ray segment really enters), otherwise a triangle could be skipped and the result would
depend on the tree.  The GPU parity tests check the end result bit for bit; this file checks the arithmetic argument itself on
the CPU (no GPU, no library): for random boxes and rays -- including rays from far away, axis-parallel rays, flat boxes and
origins on box faces -- whenever the exact (float64 / rational-free) slab test says "enters", the float32 restatements of

  * the centre / half-extent format (centerNodesKernel + FMT_CE in traceKernel), and
  * the 16-bit grid format (quantizeBvh / quantizeNodesKernel / collapseWideKernel + gridPlane arithmetic)

say "enters" too.  The restatements follow the CUDA operation order; `fma` is emulated in float64 (exact product, one rounding).
"""
import numpy as np
import pytest

f32 = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def safe_inv(d):
    d = d.astype(f32)
    tiny = np.abs(d) < f32(1e-20)
    return (f32(1.0) / np.where(tiny, np.copysign(f32(1e-20), d), d)).astype(f32)


def exact_enters(lo, hi, o, d, tmax):
    """float64 slab test of the closed box against the segment t in [0, tmax]; zero direction components handled exactly."""
    lo, hi, o, d = (x.astype(np.float64) for x in (lo, hi, o, d))
    tn = np.zeros(len(o))
    tf = np.full(len(o), float(tmax))
    ok = np.ones(len(o), bool)
    for a in range(3):
        z = d[:, a] == 0
        inside = (o[:, a] >= lo[:, a]) & (o[:, a] <= hi[:, a])
        ok &= ~z | inside
        with np.errstate(divide="ignore", invalid="ignore"):
            t0 = (lo[:, a] - o[:, a]) / d[:, a]
            t1 = (hi[:, a] - o[:, a]) / d[:, a]
        near, far = np.minimum(t0, t1), np.maximum(t0, t1)
        tn = np.where(z, tn, np.maximum(tn, near))
        tf = np.where(z, tf, np.minimum(tf, far))
    return ok & (tn <= tf)


def make_cases(n, seed):
    rng = np.random.default_rng(seed)
    ext = 20.0
    c = rng.uniform(-ext, ext, (n, 3))
    h = np.abs(rng.normal(0, 1.0, (n, 3))) * rng.choice([0.0, 1e-3, 0.05, 1.0, 5.0], (n, 1))
    h[rng.random((n, 3)) < 0.15] = 0.0                                   # flat boxes (axis-aligned quads)
    lo, hi = (c - h).astype(f32), (c + h).astype(f32)
    # rays: aimed at a point in / near the box from near, far and very far origins; some axis-parallel; some start on a face
    target = c + rng.uniform(-1.2, 1.2, (n, 3)) * (h + 1e-3)
    dist = 10.0 ** rng.uniform(-3, 7, (n, 1))
    dirn = rng.normal(size=(n, 3))
    dirn[rng.random((n, 3)) < 0.2] = 0.0
    dirn[np.abs(dirn).sum(1) == 0, 0] = 1.0
    dirn /= np.linalg.norm(dirn, axis=1, keepdims=True)
    o = target - dirn * dist
    onface = rng.random(n) < 0.1
    o[onface] = np.where(rng.random((onface.sum(), 3)) < 0.5, lo[onface], hi[onface])
    return lo, hi, o.astype(f32), dirn.astype(f32)


def ce_enters(lo, hi, o, d, best_t):
    """centerNodesKernel + the FMT_CE branch of traceKernel."""
    c = (f32(0.5) * lo + f32(0.5) * hi).astype(f32)
    e = np.maximum(hi - c, c - lo).astype(f32)
    e = ((e + (np.abs(c) + e) * f32(4.8e-7)) * f32(1.000001)).astype(f32)
    inv = safe_inv(d)
    nood = (-(o * inv)).astype(f32)
    tc = fma(c, inv, nood)
    ai = np.abs(inv)
    tn = fma(-e, ai, tc).max(axis=1)
    tf = np.minimum(fma(e, ai, tc).min(axis=1), f32(best_t))
    x = (tf * f32(1.0002)).astype(f32)
    return (tn <= x) & (x >= 0)


def grid_enters(lo, hi, o, d, best_t, root_lo, root_hi):
    """quantizeBvh grid + quantizeNodesKernel padding + the grid-space ray of traceKernel<FMT_Q32> / traceKernelWide."""
    ext = (root_hi - root_lo).astype(f32)
    max_ext = f32(max(ext.max(), 1e-30))
    ext = np.maximum(ext, f32(1e-3) * max_ext)
    cell = (ext / f32(65500.0)).astype(f32)
    origin = (root_lo - f32(8.0) * cell).astype(f32)
    ginv = (f32(1.0) / cell).astype(f32)
    qlo = np.clip(np.floor(((lo - origin) * ginv).astype(f32) - f32(1.0625)), 0, 65535).astype(f32)
    qhi = np.clip(np.ceil(((hi - origin) * ginv).astype(f32) + f32(1.0625)), 0, 65535).astype(f32)
    # ray setup
    dg = (d * ginv).astype(f32)
    inv = safe_inv(dg)
    og = ((o - origin) * ginv).astype(f32)
    a = ((f32(0.0) - og) * inv).astype(f32)
    b = ((f32(65535.0) - og) * inv).astype(f32)
    tn_c = np.maximum(np.minimum(a, b).max(axis=1), f32(0.0))
    tf_c = np.minimum(np.maximum(a, b).min(axis=1), f32(best_t))
    misses_cube = ~(tn_c <= (tf_c * f32(1.0002)).astype(f32))
    t0 = (tn_c * f32(0.99995)).astype(f32)
    og = (og + dg * t0[:, None]).astype(f32)
    nood = (t0[:, None] - ((og + f32(8388608.0)).astype(f32) * inv).astype(f32)).astype(f32)
    neg = np.signbit(inv)
    near = np.where(neg, qhi, qlo) + f32(8388608.0)          # gridPlane: 2^23 + q, exact
    far = np.where(neg, qlo, qhi) + f32(8388608.0)
    tn = np.maximum(fma(near.astype(f32), inv, nood).max(axis=1), f32(0.0))
    tf = np.minimum(fma(far.astype(f32), inv, nood).min(axis=1), f32(best_t))
    return ~misses_cube & (tn <= (tf * f32(1.0002)).astype(f32))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_centre_half_extent_box_test_is_conservative(seed):
    lo, hi, o, d = make_cases(200_000, seed)
    truth = exact_enters(lo, hi, o, d, 1e32)
    got = ce_enters(lo, hi, o, d, 1e32)
    assert truth.mean() > 0.2
    assert not (truth & ~got).any(), int((truth & ~got).sum())
    assert (got & ~truth).mean() < 0.2            # and it still culls: the padding is small


@pytest.mark.parametrize("seed", [4, 5, 6])
def test_grid_box_test_is_conservative(seed):
    lo, hi, o, d = make_cases(200_000, seed)
    root_lo, root_hi = lo.min(axis=0), hi.max(axis=0)
    # origins beyond ~1e5 scene extents leave the documented range of the grid-space ray (DESIGN: 122,000 extents)
    near_enough = np.abs(o).max(axis=1) < 1e5 * float((root_hi - root_lo).max())
    truth = exact_enters(lo, hi, o, d, 1e32) & near_enough
    got = grid_enters(lo, hi, o, d, 1e32, root_lo, root_hi)
    assert truth.mean() > 0.2
    assert not (truth & ~got).any(), int((truth & ~got).sum())


def test_best_t_clip_is_conservative():
    """With a finite far clip (a hit already found) boxes entered before it must still pass."""
    lo, hi, o, d = make_cases(100_000, 9)
    root_lo, root_hi = lo.min(axis=0), hi.max(axis=0)
    for best_t in (0.5, 30.0, 1e4):
        truth = exact_enters(lo, hi, o, d, best_t) & (np.abs(o).max(axis=1) < 1e5 * float((root_hi - root_lo).max()))
        assert not (truth & ~ce_enters(lo, hi, o, d, best_t)).any()
        assert not (truth & ~grid_enters(lo, hi, o, d, best_t, root_lo, root_hi)).any()
