"""CPU tests of the product's kernel bodies (csrc/raht_core.cuh,
raht_pipeline.cuh, pcc_arith.cuh) compiled for the host by tests/emu and run
as in-order loops, against the oracle.  This exercises stage planning, node
construction, coefficient addressing and the dataflow bookkeeping without a
GPU; the GPU tests (-m gpu) run the same bodies as CUDA kernels."""
import os

import numpy as np
import pytest

from pcc_testlib import *  # noqa


def _cmp(xyz, attrs, p, qs, qpo=None):
    mort, a_s, order = sort_cloud(xyz, attrs)
    q = qpo[order] if qpo is not None else None
    orec, ocoef = oracle_raht(1, p, qs, mort, a_s, qpoffs=q)
    erec, ecoef = emu_raht(1, p, qs, mort, a_s, qpoffs=q)
    assert np.array_equal(ecoef, ocoef)
    assert np.array_equal(erec, orec)
    erec2, _ = emu_raht(0, p, qs, mort, a_s * 0, coeffs=ocoef, qpoffs=q)
    assert np.array_equal(erec2, orec)


def test_arith_golden():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "arith_golden.npz"))
    e = load_emu()
    assert [e.emu_isqrt(int(x)) for x in g["xs"]] == g["isqrt"].tolist()
    assert [e.emu_irsqrt(int(x)) for x in g["xs"]] == g["irsqrt"].tolist()
    assert [e.emu_fixed_mul(int(a), int(b)) for a, b in zip(g["fa"], g["fb"])] == g["fxmul"].tolist()
    assert [e.emu_quantize(int(q), int(x)) for q, x in zip(g["qps"], g["qx"])] == g["quant"].tolist()
    assert [e.emu_scale(int(q), int(x)) for q, x in zip(g["qps"], g["qx"])] == g["scale"].tolist()
    assert [e.emu_morton_addr(*map(int, p)) for p in g["pts"]] == g["morton"].tolist()
    assert [e.emu_morton3d_add(int(a), int(b)) for a, b in zip(g["ma"], g["mb"])] == g["madd"].tolist()
    assert [e.emu_div_approx(int(a), int(b), 0) for a, b in zip(g["da"], g["db"])] == g["divapprox"].tolist()


@pytest.mark.parametrize("kw", [dict(), dict(prediction=0), dict(subnode=0), dict(haar=1),
                                dict(ext=0), dict(thr0=0, thr1=1)])
def test_shell(kw):
    xyz, attrs = cloud_shell(20000, bits=8, seed=3)
    for qp in (16, 34):
        _cmp(xyz, attrs, make_params(**kw), make_qpset(qp=qp))


def test_dups_lidar_sparse():
    for a in (1, 3):
        xyz, attrs = cloud_shell(20000, bits=7, seed=4, a=a, dups=True)
        for kw in (dict(), dict(haar=1), dict(ext=0)):
            _cmp(xyz, attrs, make_params(**kw), make_qpset(qp=28))
    xyz, attrs = cloud_lidar(30000, seed=2)
    _cmp(xyz, attrs, make_params(search_range=2500), make_qpset(qp=34))
    _cmp(xyz, attrs, make_params(search_range=5), make_qpset(qp=34))
    xyz, attrs = cloud_random(20000, 3, seed=12)
    _cmp(xyz, attrs, make_params(), make_qpset(qp=34))
    for bits in (4, 21):
        xyz, attrs = cloud_random(10000, bits, seed=bits, dup_frac=0.2)
        _cmp(xyz, attrs, make_params(thr0=0, thr1=1), make_qpset(qp=30))


def test_qp_structures_and_edges():
    rng = np.random.default_rng(7)
    xyz, attrs = cloud_shell(20000, bits=8, seed=5)
    qpo = rng.integers(-6, 7, size=(xyz.shape[0], 2)).astype(np.int32)
    _cmp(xyz, attrs, make_params(), make_qpset(qp=30), qpo)
    _cmp(xyz, attrs, make_params(), make_qpset(layers=[(40, -2), (36, -1), (32, 0), (28, 1)]))
    ac = [[(l - c, c - l) for c in range(7)] for l in range(4)]
    _cmp(xyz, attrs, make_params(), make_qpset(qp=30, ac_qps=ac), qpo)
    for n in (1, 2, 3, 9):
        xyz, attrs = cloud_random(n, 3, seed=n)
        _cmp(xyz, attrs, make_params(), make_qpset(qp=20))
    xyz = np.tile(np.array([[5, 6, 7]], dtype=np.int32), (6, 1))
    attrs = rng.integers(0, 256, size=(6, 3)).astype(np.int32)
    _cmp(xyz, attrs, make_params(), make_qpset(qp=20))
    xyz = np.array([[0, 0, 0], [2**20, 2**20, 2**20], [2**20 + 1, 2**20, 2**20]], dtype=np.int32)
    _cmp(xyz, attrs[:3], make_params(thr0=0, thr1=0), make_qpset(qp=20))


from test_oracle_vs_reference import LOD_CASES  # noqa: E402


@pytest.mark.parametrize("kw", LOD_CASES)
def test_lod_bodies(kw):
    """lod_core.cuh / lod_pipeline.cuh (host build) against the LoD oracle"""
    for xyz in (cloud_shell(20000, bits=8, seed=3)[0], cloud_lidar(20000, seed=2)[0],
                cloud_random(8000, 21, seed=5, dup_frac=0.1)[0], cloud_random(5000, 4, seed=6)[0]):
        lp = make_lod_params(**dict(kw, levels=8))
        op, oi, on = oracle_lod_build(lp, xyz)
        ep, ei, en = emu_lod_build(lp, xyz)
        assert np.array_equal(on, en) and np.array_equal(oi, ei) and np.array_equal(op, ep)


@pytest.mark.parametrize("a", [1, 3])
def test_lifting_coder_bodies(a):
    """lift_pipeline.cuh (host build): the whole lifting attribute coder minus
    entropy coding, encoder and decoder, against the oracle chain"""
    xyz, attrs = cloud_shell(15000, bits=8, seed=7, a=a)
    for dec, lcp, qp in ((0, 1, 34), (1, 0, 16), (2, 1, 22)):
        lp = make_lod_params(levels=8, decimation=dec)
        qs = make_qpset(qp=qp, chroma_offset=-2 if a == 3 else 0, fixed_point_qp_offset=24,
                        layers=[(qp, -2 if a == 3 else 0), (qp + 2, 0), (qp + 4, 1)])
        ov, orr, ol = oracle_lift_encode(lp, qs, lcp, xyz, attrs)
        ev, er, el = emu_attr_lift(1, lp, qs, lcp, xyz, attrs)
        assert np.array_equal(ev, ov) and np.array_equal(er, orr)
        if a == 3 and lcp:
            assert np.array_equal(el, ol)
        _, dr, _ = emu_attr_lift(0, lp, qs, lcp, xyz, attrs * 0, values=ov, lcp=ol)
        assert np.array_equal(dr, orr)


def test_lifting_coder_empty_lods():
    """empty levels of detail (equal cumulative sizes): the reference's
    per-LoD counters stop advancing; must be reproduced"""
    rng = np.random.default_rng(1)
    xyz = (rng.integers(0, 1 << 12, size=(12000, 3)) * 8).astype(np.int32)
    attrs = rng.integers(0, 256, size=(12000, 3)).astype(np.int32)
    xyz2, attrs2 = cloud_lidar(60000, seed=4, a=3)
    qs = make_qpset(qp=30, chroma_offset=-2, fixed_point_qp_offset=24, layers=[(30, -2), (32, 0), (34, 1)])
    for x, a_, d2 in ((xyz, attrs, 0), (xyz, attrs, 2), (xyz2, attrs2, 0)):
        lp = make_lod_params(levels=10, decimation=0, dist2=d2)
        ov, orr, ol = oracle_lift_encode(lp, qs, 1, x, a_)
        ev, er, el = emu_attr_lift(1, lp, qs, 1, x, a_)
        assert np.array_equal(ev, ov) and np.array_equal(er, orr) and np.array_equal(el, ol)


def test_spherical_bodies():
    """spherical.cuh (host build) against the oracle: conversion, bounding box,
    offsetAndScale, and the fused call with either offset"""
    emu, orc = load_emu(), load_oracle()
    rng = np.random.default_rng(11)
    for y, x in zip(rng.integers(-(1 << 30), 1 << 30, 3000), rng.integers(-(1 << 30), 1 << 30, 3000)):
        assert emu.emu_iatan2(int(y), int(x)) == orc.oracle_iatan2(int(y), int(x))
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "spherical_golden.npz"))
    for name in g["names"]:
        origin, theta, xyz, w = (g[f"{name}/{k}"] for k in ("origin", "theta", "xyz", "weight"))
        r, b = emu_xyz_to_rpl(origin, theta, xyz)
        assert np.array_equal(r, g[f"{name}/rpl"]) and np.array_equal(b, g[f"{name}/bbox"])
        s, b2 = emu_xyz_to_rpl(origin, theta, xyz, weight=w)  # offset = bounding-box minimum
        assert np.array_equal(s, g[f"{name}/scaled"]) and np.array_equal(b2, b)
        mp = (3, -2, 1)
        s2, _ = emu_xyz_to_rpl(origin, theta, xyz, weight=w, min_pos=mp)
        assert np.array_equal(s2, oracle_offset_and_scale(mp, w, r))
    wide = rng.integers(-(1 << 21), 1 << 21, size=(20000, 3)).astype(np.int32)
    theta = lidar_lasers(48, -1.0, 1.0)
    r, b = emu_xyz_to_rpl((9, 9, 9), theta, wide)
    o, ob = oracle_xyz_to_rpl((9, 9, 9), theta, wide)
    assert np.array_equal(r, o) and np.array_equal(b, ob)


def test_symbol_bodies():
    """symbols.cuh (host build) against the oracle: synthetic coefficient
    planes with every run-length pattern, and real transform output"""
    rng = np.random.default_rng(2)
    for a in (1, 3):
        for n, density in ((1, 1.0), (1, 0.0), (7, 0.5), (5000, 0.02), (5000, 0.9), (4096, 0.0)):
            coef = (rng.integers(-40, 41, size=(a, n)) * (rng.random((a, n)) < density)).astype(np.int32)
            o, e = oracle_coeff_symbols(coef), emu_coeff_symbols(coef)
            assert np.array_equal(o[0], e[0]) and np.array_equal(o[1], e[1]) and o[3] == e[3]
            assert (o[2] is None and e[2] is None) or np.array_equal(o[2], e[2])
            # the stream expands back to the planes
            back = np.zeros_like(coef)
            pos = np.cumsum(o[0] + 1) - 1
            back[:, pos] = o[1].T
            assert np.array_equal(back, coef) and (len(pos) == 0 or pos[-1] + o[3] == n - 1)
        xyz, attrs = cloud_shell(12000, bits=8, seed=4, a=a)
        mort, a_s, order = sort_cloud(xyz, attrs)
        _, coef = oracle_raht(1, make_params(), make_qpset(qp=28), mort, a_s)
        o, e = oracle_coeff_symbols(coef), emu_coeff_symbols(coef)
        assert all(np.array_equal(x, y) for x, y in zip(o[:2], e[:2])) and o[3] == e[3]


def test_estimate_dist2_bodies():
    """dist2.cuh (host build) against the oracle"""
    from test_oracle_vs_reference import _dist2_cases, DIST2_PARAMS

    for xyz in _dist2_cases():
        for period, rng_, pct in DIST2_PARAMS:
            assert emu_estimate_dist2(xyz, period, rng_, pct) == oracle_estimate_dist2(xyz, period, rng_, pct)


def test_quant_weight_variant_bodies():
    """lifting.cuh (host build): fixed-weight and scalable quantisation weights"""
    from test_oracle_vs_reference import _qw_structures

    for preds, npl in _qw_structures():
        for nw in ((256, 128, 64), (8192, 0, 5)):
            assert np.array_equal(emu_quant_weights_fixed(preds, npl, nw), oracle_quant_weights_fixed(preds, nw))
        n = len(preds)
        for num_points, min_log2 in ((n, 0), (3 * n + 7, 1)):
            assert np.array_equal(emu_quant_weights_scalable(npl, num_points, min_log2),
                                  oracle_quant_weights_scalable(npl, num_points, min_log2))
