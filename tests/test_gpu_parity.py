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
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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


@pytest.mark.parametrize("tex", [0, 16, 40])
def test_full_size_vs_oracle(b200, tex):
    """BASELINE config-2 size against the oracle, bit-exact: 1M-point LiDAR
    cloud, RGB and reflectance, smooth attributes (zero runs thousands of
    coefficients long) and textured ones (most positions quantise to 1 or 2:
    every block asks for the zero-run state of its predecessors)."""
    from pcc_attr_b200.synth import texture

    xyz, rgb = cloud_lidar(1000000, seed=2)
    refl = ((rgb[:, :1] * 3 + rgb[:, 1:2]) // 4).astype(np.int32)
    if tex:
        rgb, refl = texture(rgb, tex, 11), texture(refl, tex + 8, 12)
    params, qpset = make_params(search_range=2500), make_qpset(qp=34)
    p, q = _as(b200, params, qpset)
    for attrs in (rgb, refl):
        mort, a_s, order = sort_cloud(xyz, attrs)
        orec, ocoef = oracle_raht(1, params, qpset, mort, a_s)
        exp = np.empty_like(orec)
        exp[order] = np.clip(orec, 0, 255)
        rec, coef = b200.attr_raht_encode(p, q, xyz, attrs)
        assert np.array_equal(coef, ocoef), (tex, attrs.shape)
        assert np.array_equal(rec, exp), (tex, attrs.shape)
        assert np.array_equal(b200.attr_raht_decode(p, q, xyz, coef), exp)


@pytest.mark.parametrize("case", ["lidar", "shell_dups", "haar", "noext_nosub", "layers", "aclayers"])
def test_multi_attribute_pass(b200, case):
    """Colour + reflectance of a slice in ONE pass (shared sort, tree, geometry,
    dependency chain; own QpSet / coefficients / zero-run state per attribute)
    against the oracle run once per attribute, encoder and decoder."""
    from pcc_attr_b200.synth import texture

    kw, q1kw, q2kw = {}, dict(qp=34), dict(qp=28, chroma_offset=0)
    if case == "lidar":
        xyz, rgb = cloud_lidar(300000, seed=5)
        rgb = texture(rgb, 20, 3)
    elif case == "shell_dups":
        xyz, rgb = cloud_shell(60000, bits=9, seed=8, dups=True)
    else:
        xyz, rgb = cloud_shell(50000, bits=9, seed=9)
        rgb = texture(rgb, 24, 4)
    if case == "haar":
        kw = dict(haar=1)
    elif case == "noext_nosub":
        kw = dict(ext=0, subnode=0)
    elif case == "layers":
        q1kw = dict(qp=30, layers=[(30, -2), (34, 0), (38, 2)])
        q2kw = dict(qp=22)
    elif case == "aclayers":  # no fused path in the encoder: coded one by one internally
        q1kw = dict(qp=34, ac_qps=[[(c % 3 - 1, (c + 1) % 3 - 1) for c in range(7)], [(1, 0)] * 7])
    refl = texture(((rgb[:, :1] * 2 + rgb[:, 2:3]) // 3).astype(np.int32), 12, 6)
    params = make_params(search_range=500, **kw)
    qs = [make_qpset(**q1kw), make_qpset(**q2kw)]
    p = b200.RahtParams.from_buffer_copy(bytes(params))
    q = [b200.QpSet.from_buffer_copy(bytes(x)) for x in qs]
    recs, coefs = b200.attr_raht_encode_multi(p, q, xyz, [rgb, refl])
    for s, attrs in enumerate((rgb, refl)):
        mort, a_s, order = sort_cloud(xyz, attrs)
        orec, ocoef = oracle_raht(1, params, qs[s], mort, a_s)
        exp = np.empty_like(orec)
        exp[order] = np.clip(orec, 0, 255)
        assert np.array_equal(coefs[s], ocoef), (case, s)
        assert np.array_equal(recs[s], exp), (case, s)
    dec = b200.attr_raht_decode_multi(p, q, xyz, coefs)
    for s in range(2):
        assert np.array_equal(dec[s], recs[s]), (case, s)


@pytest.mark.parametrize("case", ["enc_rgb_refl", "haar", "one_set", "aclayers"])
def test_multi_batch_gang(b200, case):
    """Many units (slices / frames) in one call: the top-down passes of the
    units of a gang share their launches (k_block_warp_gang).  Units of
    different sizes and depths, against the oracle run per unit and attribute;
    decoder through the same entry."""
    from pcc_attr_b200.synth import texture

    kw, q1kw, q2kw = {}, dict(qp=34), dict(qp=28, chroma_offset=0)
    if case == "haar":
        kw = dict(haar=1)
    elif case == "aclayers":  # no fused path: unit by unit internally
        q1kw = dict(qp=34, ac_qps=[[(c % 3 - 1, (c + 1) % 3 - 1) for c in range(7)], [(1, 0)] * 7])
    units = []
    for u, (n, bits) in enumerate([(30000, 9), (2000, 7), (70000, 10), (9, 4), (45000, 9),
                                   (120000, 11), (15000, 8)]):
        if u % 2:
            xyz, rgb = cloud_lidar(n, seed=40 + u)
        else:
            xyz, rgb = cloud_shell(n, bits=bits, seed=40 + u, dups=(u == 4))
        rgb = texture(rgb, 20, 3 + u)
        refl = texture(((rgb[:, :1] * 2 + rgb[:, 2:3]) // 3).astype(np.int32), 12, 6 + u)
        units.append((xyz, [rgb, refl] if case != "one_set" else [rgb]))
    params = make_params(search_range=500, **kw)
    qs = [make_qpset(**q1kw), make_qpset(**q2kw)]
    if case == "one_set":
        qs = qs[:1]
    p = b200.RahtParams.from_buffer_copy(bytes(params))
    q = [b200.QpSet.from_buffer_copy(bytes(x)) for x in qs]
    expected = {}
    # spread over the lanes / gangs of three / gangs of four with one CTA per unit
    for gang, ctas, chain in (("0", None, None), ("3", None, None), ("4", "1", None)):
        os.environ["PCCB200_GANG"] = gang
        for k, v in (("PCCB200_GANG_CTAS", ctas),):
            os.environ.pop(k, None)
            if v:
                os.environ[k] = v
        recs, coefs = b200.attr_raht_encode_multi_batch(
            p, q, [x for x, _ in units], [a for _, a in units])
        for u, (xyz, attrs) in enumerate(units):
            for s, at in enumerate(attrs):
                if gang == "0":
                    mort, a_s, order = sort_cloud(xyz, at)
                    orec, ocoef = oracle_raht(1, params, qs[s], mort, a_s)
                    exp = np.empty_like(orec)
                    exp[order] = np.clip(orec, 0, 255)
                    expected[(u, s)] = (ocoef, exp)
                ocoef, exp = expected[(u, s)]
                assert np.array_equal(coefs[u][s], ocoef), (case, gang, ctas, chain, u, s)
                assert np.array_equal(recs[u][s], exp), (case, gang, ctas, chain, u, s)
        dec = b200.attr_raht_decode_multi_batch(p, q, [x for x, _ in units], coefs)
        for u in range(len(units)):
            for s in range(len(qs)):
                assert np.array_equal(dec[u][s], recs[u][s]), (case, gang, ctas, chain, u, s)
    os.environ.pop("PCCB200_GANG_CTAS", None)
    del os.environ["PCCB200_GANG"]


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


def test_dropin_translation_unit(b200):
    """The reference's own C++ entry points (RAHT.h:47-69), defined by the
    product's drop-in translation unit instead of tmc3/RAHT.cpp and called
    through the reference-side shim: same results as the oracle."""
    import ctypes as C

    path = os.path.join(ROOT, "oracle", "_ref", "libtmc13_dropin.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libtmc13_dropin.so not built (make -C oracle dropin)")
    lib = C.CDLL(path)
    lib.tmc13ref_raht.restype = C.c_double
    import pcc_testlib as tl

    rng = np.random.default_rng(5)
    xyz, attrs = cloud_shell(60000, bits=9, seed=77)
    qpo = rng.integers(-3, 4, size=(xyz.shape[0], 2)).astype(np.int32)
    for kw, q in ((dict(), None), (dict(haar=1), None), (dict(ext=0), None), (dict(), qpo)):
        params, qpset = make_params(**kw), make_qpset(qp=30)
        mort, a_s, order = sort_cloud(xyz, attrs)
        qq = q[order] if q is not None else None
        orec, ocoef = oracle_raht(1, params, qpset, mort, a_s, qpoffs=qq)
        drec, dcoef, _ = tl._run_raht(lib.tmc13ref_raht, 1, params, qpset, mort, a_s, None, qq)
        assert np.array_equal(dcoef, ocoef) and np.array_equal(drec, orec)
        drec2, _, _ = tl._run_raht(lib.tmc13ref_raht, 0, params, qpset, mort, a_s * 0, ocoef, qq)
        assert np.array_equal(drec2, orec)


def test_whole_codec_bitstream(b200, tmp_path):
    """The reference's tmc3 executable with RAHT.cpp replaced by the drop-in
    translation unit produces the same bitstream and reconstruction as the
    unmodified executable, and decodes the reference's bitstream identically
    (the verification recipe of scripts/Makefile.tmc13-step: md5 of bitstream,
    encoder reconstruction and decoder output)."""
    import hashlib

    import codec_harness as ch

    if not (os.path.exists(ch.REF_BIN) and os.path.exists(ch.B200_BIN)):
        pytest.skip("oracle/_ref/tmc3_{ref,b200} not built (make -C oracle codec)")
    xyz, rgb = cloud_shell(120000, bits=10, seed=9)
    ply = str(tmp_path / "in.ply")
    ch.write_ply(ply, xyz, rgb)

    def md5(p):
        return hashlib.md5(open(p, "rb").read()).hexdigest()

    for qp in (34, 22):
        out = {}
        for name, binary in (("ref", ch.REF_BIN), ("b200", ch.B200_BIN)):
            b, r = str(tmp_path / f"{name}{qp}.bin"), str(tmp_path / f"{name}{qp}_rec.ply")
            rc, log = ch.encode(binary, ply, b, r, qp=qp)
            assert rc == 0, log[-2000:]
            out[name] = (b, r)
        assert md5(out["ref"][0]) == md5(out["b200"][0]), "bitstreams differ"
        assert md5(out["ref"][1]) == md5(out["b200"][1]), "encoder reconstructions differ"
        d_ref, d_b200 = str(tmp_path / f"dref{qp}.ply"), str(tmp_path / f"db200{qp}.ply")
        assert ch.decode(ch.REF_BIN, out["ref"][0], d_ref)[0] == 0
        rc, log = ch.decode(ch.B200_BIN, out["ref"][0], d_b200)
        assert rc == 0, log[-2000:]
        assert md5(d_ref) == md5(d_b200) == md5(out["ref"][1])


@pytest.mark.parametrize("tt,decimator", [(2, 0), (2, 1), (2, 2), (1, 0)])
def test_whole_codec_bitstream_lifting(b200, tmp_path, tt, decimator):
    """The reference's tmc3 executable with AttributeLods::generate replaced by
    the drop-in translation unit (host/lod_dropin.cpp: the level-of-detail build
    runs on the GPU, the reference's own lifting / predicting loops and entropy
    coder run on its predictors) produces the same bitstream and reconstruction
    as the unmodified executable, and decodes the reference's bitstream
    identically (transformType 2 = lifting, 1 = predicting)."""
    import hashlib

    import codec_harness as ch

    if not (os.path.exists(ch.REF_BIN) and os.path.exists(ch.B200_BIN)):
        pytest.skip("oracle/_ref/tmc3_{ref,b200} not built (make -C oracle codec)")
    xyz, rgb = cloud_shell(100000, bits=10, seed=12)
    ply = str(tmp_path / "in.ply")
    ch.write_ply(ply, xyz, rgb)

    def md5(p):
        return hashlib.md5(open(p, "rb").read()).hexdigest()

    flags = ch.lod_flags(qp=34, transform_type=tt, decimator=decimator)
    out = {}
    for name, binary in (("ref", ch.REF_BIN), ("b200", ch.B200_BIN)):
        b, r = str(tmp_path / f"{name}.bin"), str(tmp_path / f"{name}_rec.ply")
        rc, log = ch.encode(binary, ply, b, r, flags=flags)
        assert rc == 0, log[-2000:]
        out[name] = (b, r)
    assert md5(out["ref"][0]) == md5(out["b200"][0]), "bitstreams differ"
    assert md5(out["ref"][1]) == md5(out["b200"][1]), "encoder reconstructions differ"
    d_ref, d_b200 = str(tmp_path / "dref.ply"), str(tmp_path / "db200.ply")
    assert ch.decode(ch.REF_BIN, out["ref"][0], d_ref)[0] == 0
    rc, log = ch.decode(ch.B200_BIN, out["ref"][0], d_b200)
    assert rc == 0, log[-2000:]
    assert md5(d_ref) == md5(d_b200) == md5(out["ref"][1])


def test_lod_handle_reuse(b200):
    """Levels of detail built once per slice and used by the colour and the
    reflectance call (AttributeEncoder::_lods): same results as the calls that
    rebuild them; the reuse rule follows AttributeLods::isReusable."""
    import ctypes as C

    xyz, rgb = cloud_shell(80000, bits=10, seed=21)
    refl = ((rgb[:, :1] + rgb[:, 1:2]) // 2).astype(np.int32)
    lp = _as_lod(b200, make_lod_params(levels=10))
    q = b200.QpSet.from_buffer_copy(bytes(make_qpset(qp=34, fixed_point_qp_offset=24)))
    h = C.c_void_p()
    L = b200.lib()
    x = np.ascontiguousarray(xyz, dtype=np.int32)
    assert L.pccb200_lod_create(C.byref(lp), x.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int32(len(x)),
                                C.byref(h)) == 0
    try:
        assert L.pccb200_lod_reusable(h, C.byref(lp)) == 1
        lp2 = _as_lod(b200, make_lod_params(levels=10, k=2))
        assert L.pccb200_lod_reusable(h, C.byref(lp2)) == 0
        for attrs, lcp in ((rgb, 1), (refl, 0)):
            val0, rec0, lcp0 = b200.attr_lift_encode(lp, q, xyz, attrs, lcp_enabled=lcp)
            a = np.ascontiguousarray(attrs, dtype=np.int32).copy()
            vals = np.empty_like(a)
            lc = np.zeros(33, dtype=np.int8)
            rc = L.pccb200_attr_lift_encode_lod(
                h, C.byref(q), C.c_int32(lcp), None, a.ctypes.data_as(C.POINTER(C.c_int32)),
                C.c_int32(a.shape[1]), C.c_int32(8), vals.ctypes.data_as(C.POINTER(C.c_int32)),
                lc.ctypes.data_as(C.POINTER(C.c_int8)))
            assert rc == 0
            assert np.array_equal(a, rec0) and np.array_equal(vals, val0)
            assert np.array_equal(lc[:10], np.asarray(lcp0)[:10])
            dec = np.empty_like(a)
            rc = L.pccb200_attr_lift_decode_lod(
                h, C.byref(q), C.c_int32(lcp), None, dec.ctypes.data_as(C.POINTER(C.c_int32)),
                C.c_int32(a.shape[1]), C.c_int32(8), vals.ctypes.data_as(C.POINTER(C.c_int32)),
                lc.ctypes.data_as(C.POINTER(C.c_int8)))
            assert rc == 0 and np.array_equal(dec, a)
    finally:
        L.pccb200_lod_destroy(h)


def _as_lod(pb, lp):
    return pb.LodParams.from_buffer_copy(bytes(lp))


def test_lod_golden(b200):
    from golden.make_golden import LOD_GOLDEN_CASES

    g = np.load(os.path.join(GOLD, "lod_golden.npz"))
    for cname in ("shell", "sparse"):
        for i, kw in enumerate(LOD_GOLDEN_CASES):
            p, idx, npl = b200.lod_build(_as_lod(b200, make_lod_params(**kw)), g[f"{cname}/xyz"])
            assert np.array_equal(npl, g[f"{cname}/{i}/npl"]), (cname, i)
            assert np.array_equal(idx, g[f"{cname}/{i}/indexes"]), (cname, i)
            assert np.array_equal(p, g[f"{cname}/{i}/preds"]), (cname, i)


@pytest.mark.parametrize("kw", [
    dict(), dict(distribution=0), dict(decimation=1), dict(decimation=2),
    dict(decimation=0, skip_layers=0, intra_range=128, inter_range=128, blending=1),
    dict(decimation=1, skip_layers=0, intra_range=16, inter_range=16, blending=1, period=3),
    dict(decimation=2, k=1, inter_range=8), dict(decimation=0, bias=(1, 2, 3), levels=6, dist2=1)])
def test_lod_vs_oracle(b200, kw):
    """LoD build on the device (Morton sort, subsampling dataflow, the atlas /
    window neighbour search, weights) against the oracle: numPointsInLod,
    indexes and every predictor bit-exact."""
    for xyz in (cloud_shell(150000, bits=10, seed=3)[0], cloud_lidar(100000, seed=2)[0],
                cloud_random(30000, 21, seed=5, dup_frac=0.1)[0], cloud_random(20000, 5, seed=6)[0]):
        lp = make_lod_params(**kw)
        op, oi, on = oracle_lod_build(lp, xyz)
        gp, gi, gn = b200.lod_build(_as_lod(b200, lp), xyz)
        assert np.array_equal(gn, on)
        assert np.array_equal(gi, oi)
        assert np.array_equal(gp, op)


def test_lifting_end_to_end(b200):
    """LoD build -> quantisation weights -> forward lifting -> inverse lifting,
    all on the device, against the oracle chain; inverse(forward(x)) == x."""
    rng = np.random.default_rng(3)
    xyz, rgb = cloud_shell(200000, bits=10, seed=12)
    lp = make_lod_params(levels=12)
    gp, gi, gn = b200.lod_build(_as_lod(b200, lp), xyz)
    op, oi, on = oracle_lod_build(lp, xyz)
    assert np.array_equal(gp, op) and np.array_equal(gi, oi) and np.array_equal(gn, on)
    qw = b200.quant_weights(gp, gn)
    assert np.array_equal(qw, oracle_quant_weights(op))
    attrs = rgb[gi].astype(np.int64) << 8
    fwd = b200.lift(True, gp, qw, gn, attrs)
    assert np.array_equal(fwd, oracle_lift(1, op, qw, on, attrs))
    inv = b200.lift(False, gp, qw, gn, fwd)
    assert np.array_equal(inv, attrs)


def test_lifting_slices_device_pointers(b200):
    """the device-pointer lifting entries (slices of a frame, coded in place)
    against the host-pointer single-slice entry, encoder and decoder"""
    import torch

    xyz, rgb = cloud_shell(90000, bits=9, seed=31)
    offs = np.array([0, 40000, 90000], dtype=np.int64)
    lp = b200.LodParams.from_buffer_copy(bytes(make_lod_params(levels=5)))
    lq = b200.QpSet.from_buffer_copy(bytes(make_qpset(qp=30, fixed_point_qp_offset=24)))
    dev = torch.device("cuda", 0)
    dxyz = torch.from_numpy(np.ascontiguousarray(xyz, dtype=np.int32)).to(dev)
    dattr = torch.from_numpy(np.ascontiguousarray(rgb, dtype=np.int32)).to(dev)
    dvals = torch.zeros_like(dattr)
    lcp = np.zeros((2, b200.MAX_LODS), dtype=np.int8)
    torch.cuda.synchronize()
    b200.attr_lift_slices_dev(True, lp, lq, 1, dxyz.data_ptr(), dattr.data_ptr(), 3, offs,
                              dvals.data_ptr(), lcp)
    rec, vals = dattr.cpu().numpy(), dvals.cpu().numpy()
    for s in range(2):
        sl = slice(int(offs[s]), int(offs[s + 1]))
        v1, r1, l1 = b200.attr_lift_encode(lp, lq, xyz[sl], rgb[sl], lcp_enabled=1)
        assert np.array_equal(vals[sl], v1) and np.array_equal(rec[sl], r1), s
        assert np.array_equal(lcp[s, :len(l1)], l1), s
    dout = torch.zeros_like(dattr)
    torch.cuda.synchronize()
    b200.attr_lift_slices_dev(False, lp, lq, 1, dxyz.data_ptr(), dout.data_ptr(), 3, offs,
                              dvals.data_ptr(), lcp)
    assert np.array_equal(dout.cpu().numpy(), rec)


@pytest.mark.parametrize("a", [1, 3])
def test_lifting_attribute_coder(b200, a):
    """pccb200_attr_lift_encode / _decode (LoD build, weights, lifting, LCP +
    quantisation, reconstruction, all on the device) against the oracle chain,
    which tests/test_oracle_vs_reference.py pins to the reference's own
    lifting encoder."""
    for xyz, attrs in (cloud_shell(120000, bits=10, seed=7, a=a), cloud_lidar(80000, seed=4, a=a)):
        for dec, lcp, qp in ((0, 1, 34), (1, 0, 16), (2, 1, 22)):
            lp = make_lod_params(levels=10, decimation=dec)
            qs = make_qpset(qp=qp, chroma_offset=-2 if a == 3 else 0, fixed_point_qp_offset=24,
                            layers=[(qp, -2 if a == 3 else 0), (qp + 2, 0), (qp + 4, 1)])
            ov, orr, ol = oracle_lift_encode(lp, qs, lcp, xyz, attrs)
            q2 = b200.QpSet.from_buffer_copy(bytes(qs))
            gv, gr, gl = b200.attr_lift_encode(_as_lod(b200, lp), q2, xyz, attrs, lcp_enabled=lcp)
            assert np.array_equal(gv, ov) and np.array_equal(gr, orr)
            if a == 3 and lcp:
                assert np.array_equal(gl, ol)
            gd = b200.attr_lift_decode(_as_lod(b200, lp), q2, xyz, ov, lcp=ol if (a == 3 and lcp) else None)
            assert np.array_equal(gd, orr)


def test_spherical_positions(b200):
    """spherical-coordinate conversion for attribute coding (row N2): golden
    vectors of the compiled reference, then a full-size LiDAR frame and an
    adversarial random cloud against the oracle"""
    pb = b200
    g = np.load(os.path.join(GOLD, "spherical_golden.npz"))
    for name in g["names"]:
        origin, theta, xyz, w = (g[f"{name}/{k}"] for k in ("origin", "theta", "xyz", "weight"))
        r, (mn, mx) = pb.xyz_to_rpl(origin, theta, xyz)
        assert np.array_equal(r, g[f"{name}/rpl"])
        assert np.array_equal(np.concatenate([mn, mx]), g[f"{name}/bbox"])
        assert np.array_equal(pb.offset_and_scale(mn, w, r), g[f"{name}/scaled"])
        s, (mn2, mx2) = pb.attr_spherical_positions(origin, theta, w, xyz)
        assert np.array_equal(s, g[f"{name}/scaled"]) and np.array_equal(mn2, mn)
        mp = (3, -2, 1)
        s2, _ = pb.attr_spherical_positions(origin, theta, w, xyz, min_pos=mp)
        assert np.array_equal(s2, oracle_offset_and_scale(mp, w, r))
    rng = np.random.default_rng(3)
    xyz, _ = cloud_lidar(1000000, seed=2)
    wide = rng.integers(-(1 << 21), 1 << 21, size=(300000, 3)).astype(np.int32)
    for pts, origin, theta in ((xyz, (40, -25, 310), lidar_lasers(64)),
                               (wide, (0, 0, 0), lidar_lasers(48, -1.0, 1.0)),
                               (wide[:1], (0, 0, 0), lidar_lasers(1))):
        r, (mn, mx) = pb.xyz_to_rpl(origin, theta, pts)
        o, ob = oracle_xyz_to_rpl(origin, theta, pts)
        assert np.array_equal(r, o) and np.array_equal(np.concatenate([mn, mx]), ob)
    with pytest.raises(pb.PccB200Error):
        pb.xyz_to_rpl((0, 0, 0), np.zeros(0, dtype=np.int32), wide[:4])


def test_repeatability(b200):
    """the dataflow kernels are timing dependent inside (tickets, polls): the
    same call must give the same bits every time, also with a tiny persistent
    grid (looping warps) — run as a subprocess so the environment knob applies"""
    import subprocess
    import sys

    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import pcc_attr_b200 as pb\n"
        "from pcc_testlib import cloud_lidar, make_params, make_qpset\n"
        "xyz, attrs = cloud_lidar(120000, seed=6, a=3)\n"
        "p = pb.RahtParams.from_buffer_copy(bytes(make_params())); q = pb.QpSet.from_buffer_copy(bytes(make_qpset(qp=34)))\n"
        "ref = None\n"
        "for i in range(25):\n"
        "    rec, coef = pb.attr_raht_encode(p, q, xyz, attrs)\n"
        "    if ref is None: ref = (rec.copy(), coef.copy())\n"
        "    assert np.array_equal(rec, ref[0]) and np.array_equal(coef, ref[1]), i\n"
        "print('same', int(np.abs(coef).sum()))\n"
    ) % (os.path.join(ROOT, "mpeg-pcc-tmc13_b200"), os.path.join(ROOT, "tests"))
    outs = []
    for grid in ("", "8"):
        env = dict(os.environ)
        if grid:
            env["PCCB200_BLOCK_GRID"] = grid
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip())
    assert outs[0] == outs[1] and outs[0].startswith("same")


def test_entropy_coder_symbols(b200):
    """symbol preparation for the entropy coder (row N1): the stream the fused
    encoder call hands over equals the one decoded from the reference
    encoder's payload (golden), and the oracle's on larger clouds"""
    from golden.make_golden import SYMBOL_GOLDEN_CASES

    pb = b200
    g = np.load(os.path.join(GOLD, "symbols_golden.npz"))
    for name, a, qp in SYMBOL_GOLDEN_CASES:
        p, q = _as(pb, make_params(), make_qpset(qp=qp))
        rec, runs, vals, ctx, tail = pb.attr_raht_encode_symbols(p, q, g[f"{name}/xyz"], g[f"{name}/attrs"])
        assert np.array_equal(runs, g[f"{name}/runs"]) and np.array_equal(vals, g[f"{name}/values"])
        assert tail == int(g[f"{name}/tail"]) and np.array_equal(rec, g[f"{name}/recon"])
    rng = np.random.default_rng(8)
    for a in (1, 3):
        xyz, attrs = cloud_lidar(300000, seed=3, a=a)
        params, qs = make_params(), make_qpset(qp=34)
        p, q = _as(pb, params, qs)
        rec, coef = pb.attr_raht_encode(p, q, xyz, attrs)
        o = oracle_coeff_symbols(coef)
        s = pb.coeff_symbols(coef)
        rec2, runs, vals, ctx, tail = pb.attr_raht_encode_symbols(p, q, xyz, attrs)
        for got in (s, (runs, vals, ctx, tail)):
            assert np.array_equal(got[0], o[0]) and np.array_equal(got[1], o[1]) and got[3] == o[3]
            assert (o[2] is None and got[2] is None) or np.array_equal(got[2], o[2])
        assert np.array_equal(rec2, rec)
        for n, density in ((1, 0.0), (1, 1.0), (100000, 0.0), (100000, 0.7)):
            c = (rng.integers(-9, 10, size=(a, n)) * (rng.random((a, n)) < density)).astype(np.int32)
            o, s = oracle_coeff_symbols(c), pb.coeff_symbols(c)
            assert np.array_equal(s[0], o[0]) and np.array_equal(s[1], o[1]) and s[3] == o[3]


def test_estimate_dist2(b200):
    """estimateDist2 (row N3, first half) against the oracle, bench-size frame included"""
    from test_oracle_vs_reference import _dist2_cases, DIST2_PARAMS

    pb = b200
    cases = _dist2_cases() + [cloud_lidar(1000000, seed=2)[0]]
    for xyz in cases:
        for period, rng_, pct in DIST2_PARAMS:
            assert pb.estimate_dist2(xyz, period, rng_, pct) == oracle_estimate_dist2(xyz, period, rng_, pct)
    with pytest.raises(pb.PccB200Error):
        pb.estimate_dist2(cases[0], 100, 128, 1.0)


def test_quant_weight_variants(b200):
    """computeQuantizationWeights (predicting transform, fixed per-slot weights;
    levels of detail that reference themselves run the ordered body) and
    computeQuantizationWeightsScalable, against the oracle"""
    from test_oracle_vs_reference import _qw_structures

    pb = b200
    structs = _qw_structures()
    xyz, _ = cloud_shell(300000, bits=10, seed=12)
    p, idx, npl = oracle_lod_build(make_lod_params(levels=10), xyz)
    structs.append((p, npl))
    for preds, npl in structs:
        for nw in ((256, 128, 64), (8192, 0, 5)):
            assert np.array_equal(pb.quant_weights_fixed(preds, npl, nw), oracle_quant_weights_fixed(preds, nw))
        n = len(preds)
        for num_points, min_log2 in ((n, 0), (3 * n + 7, 1)):
            assert np.array_equal(pb.quant_weights_scalable(npl, num_points, min_log2),
                                  oracle_quant_weights_scalable(npl, num_points, min_log2))


def test_fuzz_slice_gpu(b200):
    """a seeded slice of tools/fuzz_gpu.py: the CUDA path against the oracle on
    random points of the parameter space (transform, attribute level, symbol
    stream, LoD build, lifting coder)"""
    import subprocess
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_gpu.py"), "90", "50", "7"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "raht 90 cases, 0 mismatches; lod/lifting 50 cases, 0 mismatches" in r.stdout
