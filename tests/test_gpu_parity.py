"""GPU parity tests (run on a B200 with `pytest -m gpu`): the CUDA path,
called through the C ABI, against the oracle and the committed golden
vectors.  Bit-exact: every comparison is np.array_equal."""
import os

import numpy as np
import pytest

from pcc_testlib import *  # noqa
from golden.make_golden import VARIANTS  # noqa

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def b200():
    import pcc_attr_b200 as pb

    pb.lib()
    return pb


def _as(pb, params, qpset):
    """testlib PODs -> binding PODs (same layout)"""
    import ctypes as C

    p = pb.RahtParams.from_buffer_copy(bytes(params))
    q = pb.QpSet.from_buffer_copy(bytes(qpset))
    return p, q


def _cmp_oracle(pb, xyz, attrs, params, qpset, qpo=None):
    mort, a_s, order = sort_cloud(xyz, attrs)
    q = qpo[order] if qpo is not None else None
    orec, ocoef = oracle_raht(1, params, qpset, mort, a_s, qpoffs=q)
    p2, q2 = _as(pb, params, qpset)
    grec, gcoef = pb.raht_forward(p2, q2, mort, a_s, qpoffs=q)
    assert np.array_equal(gcoef, ocoef)
    assert np.array_equal(grec, orec)
    grec2 = pb.raht_inverse(p2, q2, mort, ocoef, qpoffs=q)
    assert np.array_equal(grec2, orec)


def test_raht_golden(b200):
    g = np.load(os.path.join(GOLD, "raht_golden.npz"))
    for cname in ["cube", "shell", "shelldup", "lidar", "sparse21"]:
        xyz, attrs = g[f"{cname}/xyz"], g[f"{cname}/attrs"]
        qpo = g[f"{cname}/qpo"] if f"{cname}/qpo" in g else None
        mort, a_s, order = sort_cloud(xyz, attrs)
        q = qpo[order] if qpo is not None else None
        for vname, kw in VARIANTS.items():
            for qp in (16, 34):
                p, qs = _as(b200, make_params(**kw), make_qpset(qp=qp))
                rec, coef = b200.raht_forward(p, qs, mort, a_s, qpoffs=q)
                assert np.array_equal(coef, g[f"{cname}/{vname}/qp{qp}/coef"]), (cname, vname, qp)
                assert np.array_equal(rec, g[f"{cname}/{vname}/qp{qp}/rec"]), (cname, vname, qp)
                rec2 = b200.raht_inverse(p, qs, mort, coef, qpoffs=q)
                assert np.array_equal(rec2, rec), (cname, vname, qp)


@pytest.mark.parametrize("kw", [dict(), dict(prediction=0), dict(subnode=0), dict(haar=1),
                                dict(ext=0), dict(thr0=0, thr1=1)])
@pytest.mark.parametrize("qp", [10, 34, 46])
def test_raht_vs_oracle_shell(b200, kw, qp):
    xyz, attrs = cloud_shell(60000, bits=9, seed=qp)
    _cmp_oracle(b200, xyz, attrs, make_params(**kw), make_qpset(qp=qp))


@pytest.mark.parametrize("a", [1, 3])
def test_raht_dups_lidar(b200, a):
    xyz, attrs = cloud_shell(40000, bits=7, seed=4, a=a, dups=True)
    for kw in (dict(), dict(haar=1), dict(ext=0)):
        _cmp_oracle(b200, xyz, attrs, make_params(**kw), make_qpset(qp=28))
    xyz, attrs = cloud_lidar(100000, seed=2, a=a)
    _cmp_oracle(b200, xyz, attrs, make_params(search_range=2500), make_qpset(qp=34))
    xyz, attrs = cloud_random(30000, 3, seed=12, a=a)  # node weights > 1024
    _cmp_oracle(b200, xyz, attrs, make_params(), make_qpset(qp=34))


def test_raht_qp_structures(b200):
    rng = np.random.default_rng(7)
    xyz, attrs = cloud_shell(50000, bits=9, seed=5)
    qpo = rng.integers(-6, 7, size=(xyz.shape[0], 2)).astype(np.int32)
    _cmp_oracle(b200, xyz, attrs, make_params(), make_qpset(qp=30), qpo)
    _cmp_oracle(b200, xyz, attrs, make_params(),
                make_qpset(layers=[(40, -2), (36, -1), (32, 0), (28, 1), (26, 2)]))
    ac = [[(l - c, c - l) for c in range(7)] for l in range(4)]
    _cmp_oracle(b200, xyz, attrs, make_params(), make_qpset(qp=30, ac_qps=ac), qpo)
    xyz, a16 = cloud_random(30000, 8, seed=11, bitdepth=16)
    _cmp_oracle(b200, xyz, a16, make_params(), make_qpset(qp=40, bitdepth=16))


def test_raht_edge_cases(b200):
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 9, 17, 257):
        xyz, attrs = cloud_random(n, 3, seed=n)
        _cmp_oracle(b200, xyz, attrs, make_params(), make_qpset(qp=20))
    xyz = np.tile(np.array([[5, 6, 7]], dtype=np.int32), (6, 1))
    attrs = rng.integers(0, 256, size=(6, 3)).astype(np.int32)
    _cmp_oracle(b200, xyz, attrs, make_params(), make_qpset(qp=20))
    xyz = np.array([[0, 0, 0], [2**20, 2**20, 2**20], [2**20 + 1, 2**20, 2**20]], dtype=np.int32)
    _cmp_oracle(b200, xyz, attrs[:3], make_params(thr0=0, thr1=0), make_qpset(qp=20))
    for bits in (4, 12, 21):
        xyz, attrs = cloud_random(20000, bits, seed=bits, dup_frac=0.2)
        _cmp_oracle(b200, xyz, attrs, make_params(thr0=0, thr1=1), make_qpset(qp=30))


def test_errors(b200):
    p, q = _as(b200, make_params(), make_qpset())
    mort = np.array([5, 3, 9], dtype=np.int64)
    with pytest.raises(b200.PccB200Error):
        b200.raht_forward(p, q, mort, np.zeros((3, 3), dtype=np.int32))


def test_morton_sort(b200):
    rng = np.random.default_rng(1)
    for n, bits in ((1, 5), (1000, 4), (4096, 10), (4097, 10), (300000, 21), (100000, 1)):
        xyz = rng.integers(0, 1 << bits, size=(n, 3), dtype=np.int32)
        keys, order = b200.morton_sort(xyz)
        ek, eo = oracle_morton_sort(xyz)
        assert np.array_equal(keys, ek)
        assert np.array_equal(order, eo)


def test_attr_level_roundtrip(b200):
    """sort + gather + transform + clip on the device == oracle pipeline;
    decoder reproduces the encoder's reconstruction."""
    xyz, attrs = cloud_lidar(200000, seed=21)
    params, qpset = make_params(search_range=2500), make_qpset(qp=34)
    p, q = _as(b200, params, qpset)
    rec, coef = b200.attr_raht_encode(p, q, xyz, attrs, bitdepth=8)
    mort, a_s, order = sort_cloud(xyz, attrs)
    orec, ocoef = oracle_raht(1, params, qpset, mort, a_s)
    exp = np.empty_like(orec)
    exp[order] = np.clip(orec, 0, 255)
    assert np.array_equal(coef, ocoef)
    assert np.array_equal(rec, exp)
    dec = b200.attr_raht_decode(p, q, xyz, coef, bitdepth=8)
    assert np.array_equal(dec, rec)


def test_slices(b200):
    """independent slices in one call == one call per slice"""
    xyz, attrs = cloud_shell(90000, bits=9, seed=8)
    offs = np.array([0, 20000, 55000, 90000], dtype=np.int64)
    p, q = _as(b200, make_params(), make_qpset(qp=30))
    rec, coef = b200.attr_raht_encode(p, q, xyz, attrs, slice_offsets=offs)
    for s in range(3):
        a, b = offs[s], offs[s + 1]
        r1, c1 = b200.attr_raht_encode(p, q, xyz[a:b], attrs[a:b])
        assert np.array_equal(rec[a:b], r1)
        assert np.array_equal(coef[:, a:b], c1)


def test_full_size_properties(b200):
    """BASELINE config-2 size: 1M-point LiDAR cloud, RGB.  Size-independent
    properties: enc -> dec reproduces the reconstruction bit-exactly, the
    coefficient count is exact, launches are counted."""
    xyz, attrs = cloud_lidar(1000000, seed=2)
    p, q = _as(b200, make_params(search_range=2500), make_qpset(qp=34))
    before = b200.kernel_launch_count()
    rec, coef = b200.attr_raht_encode(p, q, xyz, attrs)
    assert b200.kernel_launch_count() > before
    dec = b200.attr_raht_decode(p, q, xyz, coef)
    assert np.array_equal(dec, rec)
    assert coef.shape == (3, xyz.shape[0])
    assert np.abs(rec - attrs).mean() < 16  # lossy but sane at qp 34


@pytest.mark.parametrize("a", [1, 3])
def test_lifting_vs_oracle(b200, a):
    """quantisation weights and forward / inverse lifting (64-bit atomics per
    LoD) against the sequential oracle, bit-exact; inverse(forward(x)) is
    checked against the oracle's inverse as well."""
    rng = np.random.default_rng(23)
    for n, lods in ((5000, 6), (300000, 12), (37, 3), (1000000, 3)):
        preds, npl = synth_predictors(n, lods, seed=n + a)
        qw_o = oracle_quant_weights(preds)
        qw_g = b200.quant_weights(preds, npl)
        assert np.array_equal(qw_g, qw_o)
        attrs = (rng.integers(0, 256, size=(n, a)).astype(np.int64)) << 8
        fo = oracle_lift(1, preds, qw_o, npl, attrs)
        fg = b200.lift(True, preds, qw_g, npl, attrs)
        assert np.array_equal(fg, fo)
        io = oracle_lift(0, preds, qw_o, npl, fo)
        ig = b200.lift(False, preds, qw_g, npl, fg)
        assert np.array_equal(ig, io)


def test_concurrent_calls(b200):
    """calls from several host threads run on separate lanes and stay exact"""
    from concurrent.futures import ThreadPoolExecutor

    xyz, attrs = cloud_shell(80000, bits=9, seed=31)
    params, qpset = make_params(), make_qpset(qp=34)
    p, q = _as(b200, params, qpset)
    ref_rec, ref_coef = b200.attr_raht_encode(p, q, xyz, attrs)
    with ThreadPoolExecutor(max_workers=12) as pool:
        outs = list(pool.map(lambda _: b200.attr_raht_encode(p, q, xyz, attrs), range(24)))
    for rec, coef in outs:
        assert np.array_equal(rec, ref_rec) and np.array_equal(coef, ref_coef)
