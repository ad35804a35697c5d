"""CPU tests of the drop-in boundary: the library loads without a GPU, exports
exactly the symbols include/pcc_attr_b200.h declares, and refuses to compute
(loudly) when no sm_100 device is present -- there is no CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "pcc_attr_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pccb200_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported():
    import pcc_attr_b200 as pb

    lib = pb.lib()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(pb.EXPORTS) == names
    assert lib.pccb200_abi_version() == 1


def test_pod_layout_matches_header():
    import pcc_attr_b200 as pb
    from pcc_testlib import QpSet, RahtParams

    assert C.sizeof(pb.RahtParams) == 4 * (6 + 19 + 12 + 1) == C.sizeof(RahtParams)
    assert C.sizeof(pb.QpSet) == 4 * (1 + 64 + 3 + 32 * 14) == C.sizeof(QpSet)
    assert C.sizeof(pb.Predictor) == 28 == pb.PREDICTOR_DTYPE.itemsize


def test_defaults_need_no_device():
    import pcc_attr_b200 as pb

    p = pb.default_params()
    assert (p.prediction_enabled, p.prediction_threshold0, p.prediction_threshold1) == (1, 2, 6)
    assert list(p.pred_weight_parent)[:4] == [9, 3, 3, 3]
    assert list(p.pred_weight_child) == [2, 2, 5, 2, 5, 5, 2, 2, 2, 2, 2, 2]


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import pcc_attr_b200 as pb
    from pcc_testlib import make_qpset

    q = pb.QpSet.from_buffer_copy(bytes(make_qpset()))
    with pytest.raises(pb.PccB200Error):
        pb.raht_forward(pb.default_params(), q, np.arange(4, dtype=np.int64),
                        np.zeros((4, 3), dtype=np.int32))


def test_batch_entry_argument_checks():
    """The many-units entry validates its arguments before it looks for a
    device: nulls, empty units and more than four components are refused with
    PCCB200_ERR_INVALID_ARG; a well-formed call without a GPU fails loudly."""
    import torch

    import pcc_attr_b200 as pb
    from pcc_testlib import make_qpset

    q = pb.QpSet.from_buffer_copy(bytes(make_qpset()))
    p = pb.default_params()
    xyz = [np.zeros((8, 3), dtype=np.int32), np.zeros((5, 3), dtype=np.int32)]
    rgb = [np.zeros((8, 3), dtype=np.int32), np.zeros((5, 3), dtype=np.int32)]
    lib = pb.lib()
    k, m = 1, 2
    QP = C.POINTER(pb.QpSet) * k
    VP = C.c_void_p * m
    coefs = [np.zeros((3, 8), dtype=np.int32), np.zeros((3, 5), dtype=np.int32)]

    def call(xp, ap, cp, na, ns):
        return lib.pccb200_attr_raht_encode_multi_batch(
            C.byref(p), C.c_int32(k), QP(C.pointer(q)), C.c_int32(m), xp, ap,
            (C.c_int32 * k)(*na), (C.c_int32 * k)(8), (C.c_int32 * m)(*ns), cp)

    good = (VP(*[x.ctypes.data for x in xyz]), VP(*[a.ctypes.data for a in rgb]),
            VP(*[c.ctypes.data for c in coefs]))
    invalid = 1  # PCCB200_ERR_INVALID_ARG
    assert call(None, good[1], good[2], [3], [8, 5]) == invalid
    assert call(VP(xyz[0].ctypes.data, None), good[1], good[2], [3], [8, 5]) == invalid
    assert call(good[0], good[1], good[2], [3], [8, 0]) == invalid
    assert call(good[0], good[1], good[2], [5], [8, 5]) == invalid
    if not torch.cuda.is_available():
        rc = call(good[0], good[1], good[2], [3], [8, 5])
        assert rc != 0 and rc != invalid
        assert b"CUDA" in lib.pccb200_last_error() or b"device" in lib.pccb200_last_error()


def test_recolour_entry_argument_checks():
    """pccb200_recolour refuses nulls and bad component counts before it looks for
    a device; a well-formed call without a GPU fails loudly (no CPU fallback)"""
    import torch

    import pcc_attr_b200 as pb

    lib = pb.lib()
    p = pb.default_recolour_params()
    assert (p.num_neighbours_fwd, p.num_neighbours_bwd, p.search_range) == (8, 1, 1)
    assert p.dist_offset_fwd == 4.0 and p.max_geometry_dist2_bwd == 1000.0
    xyz = np.zeros((16, 3), dtype=np.int32)
    rgb = np.zeros((16, 3), dtype=np.int32)
    out = np.zeros((16, 3), dtype=np.int32)
    off = (C.c_int32 * 3)(0, 0, 0)

    def call(sx, sa, a, tx, o):
        return lib.pccb200_recolour(C.byref(p), sx, sa, C.c_int32(a), C.c_int32(16), C.c_double(1.0),
                                    off, tx, C.c_int32(16), C.c_int32(8), o)

    P = lambda x: x.ctypes.data_as(C.POINTER(C.c_int32))
    assert call(None, P(rgb), 3, P(xyz), P(out)) == 1
    assert call(P(xyz), P(rgb), 2, P(xyz), P(out)) == 1
    assert call(P(xyz), P(rgb), 3, P(xyz), None) == 1
    if not torch.cuda.is_available():
        assert call(P(xyz), P(rgb), 3, P(xyz), P(out)) not in (0, 1)
