"""Recolouring (attribute transfer to the coded geometry):
the C oracle against the compiled reference (tolerance: the two differ only in
how distance ties are broken), the product's kernel bodies run on the host
against the oracle (bit-exact), and -- on a GPU -- the CUDA path against the
oracle (bit-exact)."""
import numpy as np
import pytest

import ctypes as C

from pcc_testlib import *  # noqa
from pcc_testlib import _ptr
from pcc_attr_b200.synth import cloud_shell, texture

CASES = {
    # name: (A, scale, offset, params)
    "colour_half": (3, 0.5, (0, 0, 0), {}),
    "colour_same": (3, 1.0, (0, 0, 0), {}),
    "colour_offset": (3, 0.37, (5, 3, 9), dict(search_range=2)),
    "refl_half": (1, 0.5, (0, 0, 0), {}),
    "refl_k": (1, 0.25, (0, 0, 0), dict(num_neighbours_fwd=5, num_neighbours_bwd=2)),
    "plain_avg": (3, 0.5, (0, 0, 0), dict(use_dist_weighted_avg_fwd=0, use_dist_weighted_avg_bwd=0,
                                           skip_avg_if_identical_source_point_present_bwd=1)),
    "attr_prune": (3, 0.5, (0, 0, 0), dict(max_attribute_dist2_fwd=300., max_attribute_dist2_bwd=200.)),
    "geom_limit": (3, 0.5, (0, 0, 0), dict(max_geometry_dist2_fwd=6., max_geometry_dist2_bwd=3.)),
}


def _case(name, n=6000, bits=8, seed=11):
    a, scale, off, kw = CASES[name]
    xyz, rgb = cloud_shell(n, bits=bits, seed=seed)
    rgb = texture(rgb, 20, seed + 1)
    attrs = rgb if a == 3 else rgb[:, :1].copy()
    tgt = coded_geometry(xyz, scale)
    # posInTgt = posInSrc * scale - offset
    tgt = tgt - np.array(off, dtype=np.int32)
    keep = (tgt >= 0).all(axis=1)
    return xyz, attrs, scale, off, np.ascontiguousarray(tgt[keep]), make_recolour_params(**kw)


@pytest.mark.parametrize("name", list(CASES))
def test_kernel_bodies_vs_oracle(name):
    """the product's functors, run as in-order loops on the host (tests/emu),
    bit-exact against the brute-force oracle"""
    sx, sa, scale, off, tx, p = _case(name)
    o = oracle_recolour(p, sx, sa, scale, off, tx)
    e = emu_recolour(p, sx, sa, scale, off, tx)
    assert np.array_equal(e, o), (name, int((e != o).any(axis=1).sum()))


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_vs_reference(name):
    """The oracle against the compiled reference (nanoflann kd-trees).  The two
    agree wherever no distance tie reaches the k-th neighbour (next test: exact
    on a cloud in general position; `colour_same`: exact); nanoflann keeps
    whichever equidistant candidate its traversal met first, the oracle the one
    with the lower index -- both are k nearest neighbours.  On voxelised clouds
    at scale 1/2 or 1/4 ties are the rule (integer grids), and a different,
    equally near neighbour carries a different texture sample (+-20 here), so
    the comparison is a tolerance.  Measured: 71 - 100 % of the target points
    identical, mean absolute difference 0 - 1.5 levels, largest 45 levels."""
    if not recolourref_available():
        pytest.skip("compiled reference not present")
    sx, sa, scale, off, tx, p = _case(name)
    r = ref_recolour(p, sx, sa, scale, off, tx)
    o = oracle_recolour(p, sx, sa, scale, off, tx)
    same = (r == o).all(axis=1).mean()
    diff = np.abs(r.astype(np.int64) - o)
    assert same >= (1.0 if name == "colour_same" else 0.68), (name, same)
    assert diff.max() <= 50, (name, int(diff.max()))
    assert diff.mean() < 1.6, (name, float(diff.mean()))


def test_oracle_vs_reference_smooth_field():
    """the same comparison on an untextured attribute field: whichever of the
    equidistant neighbours is taken, the transferred value moves with the local
    gradient of the field only (measured: mean 0.34 levels, largest 16 on this
    sparse 8-bit shell)"""
    if not recolourref_available():
        pytest.skip("compiled reference not present")
    xyz, rgb = cloud_shell(6000, bits=8, seed=11)
    tx = coded_geometry(xyz, 0.5)
    p = make_recolour_params()
    r = ref_recolour(p, xyz, rgb, 0.5, (0, 0, 0), tx)
    o = oracle_recolour(p, xyz, rgb, 0.5, (0, 0, 0), tx)
    diff = np.abs(r.astype(np.int64) - o)
    assert diff.max() <= 20 and diff.mean() < 0.5, (int(diff.max()), float(diff.mean()))


def test_reference_exact_without_ties():
    """on a cloud in general position (distinct irrational-ish distances: a
    non-unit scale and jittered coordinates) no tie reaches the k-th neighbour
    and the oracle reproduces the reference exactly"""
    if not recolourref_available():
        pytest.skip("compiled reference not present")
    rng = np.random.default_rng(5)
    sx = rng.integers(0, 4000, size=(5000, 3)).astype(np.int32)
    sx = np.unique(sx, axis=0)
    sa = rng.integers(0, 256, size=(sx.shape[0], 3)).astype(np.int32)
    tx = np.unique(rng.integers(0, 1500, size=(3000, 3)).astype(np.int32), axis=0)
    p = make_recolour_params()
    r = ref_recolour(p, sx, sa, 0.3718, (3, 1, 2), tx)
    o = oracle_recolour(p, sx, sa, 0.3718, (3, 1, 2), tx)
    assert np.array_equal(r, o), int((r != o).any(axis=1).sum())


def test_identity_transfer():
    """source == target, scale 1: every point finds itself at distance 0 in both
    directions and keeps its attributes (a size-independent property)"""
    xyz, rgb = cloud_shell(20000, bits=9, seed=3)
    p = make_recolour_params()
    e = emu_recolour(p, xyz, rgb, 1.0, (0, 0, 0), xyz)
    assert np.array_equal(e, rgb)


def test_argument_checks():
    xyz, rgb = cloud_shell(100, bits=6, seed=1)
    p = make_recolour_params(num_neighbours_fwd=17)
    lib = load_emu()
    out = np.zeros_like(rgb)
    off = np.zeros(3, dtype=np.int32)
    rc = lib.emu_recolour(C.byref(p), _ptr(xyz, C.c_int32), _ptr(rgb, C.c_int32), 3, 100,
                          C.c_double(1.0), _ptr(off, C.c_int32), _ptr(xyz, C.c_int32), 100, 8,
                          _ptr(out, C.c_int32))
    assert rc != 0
    neg = xyz.copy()
    neg[0, 0] = -1
    p = make_recolour_params()
    rc = lib.emu_recolour(C.byref(p), _ptr(neg, C.c_int32), _ptr(rgb, C.c_int32), 3, 100,
                          C.c_double(1.0), _ptr(off, C.c_int32), _ptr(xyz, C.c_int32), 100, 8,
                          _ptr(out, C.c_int32))
    assert rc != 0


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_vs_oracle(name):
    import pcc_attr_b200 as pb

    sx, sa, scale, off, tx, p = _case(name, n=30000, bits=9, seed=21)
    o = oracle_recolour(p, sx, sa, scale, off, tx)
    g = pb.recolour(pb.RecolourParams.from_buffer_copy(bytes(p)), sx, sa, tx, scale, off)
    assert np.array_equal(g, o), (name, int((g != o).any(axis=1).sum()))


@pytest.mark.gpu
def test_gpu_full_size():
    """1M-point source, half-resolution target: the CUDA path against its own
    kernel bodies run on the host, and the identity property at full size"""
    import pcc_attr_b200 as pb

    xyz, rgb = cloud_shell(1000000, bits=11, seed=7)
    rgb = texture(rgb, 16, 8)
    p = make_recolour_params()
    pp = pb.RecolourParams.from_buffer_copy(bytes(p))
    assert np.array_equal(pb.recolour(pp, xyz, rgb, xyz, 1.0), rgb)
    tx = coded_geometry(xyz, 0.5)
    g = pb.recolour(pp, xyz, rgb, tx, 0.5)
    e = emu_recolour(p, xyz, rgb, 0.5, (0, 0, 0), tx)
    assert np.array_equal(g, e)


@pytest.mark.parametrize("seed", range(12))
def test_kernel_bodies_fuzz(seed):
    """random parameter sets, scales, offsets and cloud shapes (including
    isolated far points: queries outside the occupied box, rings that grow, the
    scan-everything fallback): kernel bodies on the host == oracle, bit-exact"""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(200, 3000))
    if seed % 3 == 0:
        sx = np.unique(rng.integers(0, int(rng.integers(8, 3000)), size=(n, 3)).astype(np.int32), axis=0)
    else:
        sx, _ = cloud_shell(n, bits=int(rng.integers(5, 10)), seed=seed)
    if seed % 4 == 1:  # a few outliers far away from everything
        sx = np.concatenate([sx, rng.integers(100000, 2000000, size=(5, 3)).astype(np.int32)])
    a = 3 if seed % 2 else 1
    sa = rng.integers(0, 1 << 8, size=(sx.shape[0], a)).astype(np.int32)
    scale = float(rng.choice([1.0, 0.5, 0.25, 0.731, 1.37, 2.0]))
    off = tuple(int(v) for v in rng.integers(0, 7, size=3))
    tx = coded_geometry(sx, scale) - np.array(off, dtype=np.int32)
    tx = np.ascontiguousarray(tx[(tx >= 0).all(axis=1) & (tx < (1 << 21)).all(axis=1)])
    if tx.shape[0] < 20:
        pytest.skip("degenerate target")
    kf = int(rng.integers(1, min(16, sx.shape[0]) + 1))
    kb = int(rng.integers(1, min(4, tx.shape[0]) + 1))
    p = make_recolour_params(
        num_neighbours_fwd=kf, num_neighbours_bwd=kb, search_range=int(rng.integers(0, 3)),
        use_dist_weighted_avg_fwd=int(rng.integers(0, 2)), use_dist_weighted_avg_bwd=int(rng.integers(0, 2)),
        skip_avg_if_identical_source_point_present_fwd=int(rng.integers(0, 2)),
        skip_avg_if_identical_source_point_present_bwd=int(rng.integers(0, 2)),
        max_geometry_dist2_fwd=float(rng.choice([1000., 50., 4.])),
        max_geometry_dist2_bwd=float(rng.choice([1000., 9., 2.])),
        max_attribute_dist2_fwd=float(rng.choice([1000., 400., 60.])),
        max_attribute_dist2_bwd=float(rng.choice([1000., 300.])),
        dist_offset_fwd=float(rng.choice([4., 1., 0.5])), dist_offset_bwd=float(rng.choice([4., 2.])))
    o = oracle_recolour(p, sx, sa, scale, off, tx)
    e = emu_recolour(p, sx, sa, scale, off, tx)
    assert np.array_equal(e, o), (seed, int((e != o).any(axis=1).sum()))
