"""ctypes binding of the C ABI in include/pcc_attr_b200.h (used by the tests
and bench.py; the product itself is the C ABI + CUDA kernels).

There is no CPU fallback: importing works anywhere, but every compute call
raises if libpcc_attr_b200.so is missing or no sm_100 device is present."""
import ctypes as C
import os

import numpy as np

# one hardware queue per library lane (see the header: the application sets it,
# before its CUDA context exists; this binding is the application here)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libpcc_attr_b200.so")

MAX_QP_LAYERS = 32
MAX_AC_QP_LAYERS = 32


class RahtParams(C.Structure):
    """pccb200_raht_params <- pcc::RahtPredictionParams (tmc3/hls.h:439-466)"""
    _fields_ = [
        ("prediction_enabled", C.c_int32),
        ("integer_haar", C.c_int32),
        ("prediction_threshold0", C.c_int32),
        ("prediction_threshold1", C.c_int32),
        ("subnode_prediction_enabled", C.c_int32),
        ("prediction_search_range", C.c_int32),
        ("pred_weight_parent", C.c_int32 * 19),
        ("pred_weight_child", C.c_int32 * 12),
        ("raht_extension", C.c_int32),
    ]


class QpSet(C.Structure):
    """pccb200_qpset <- pcc::QpSet (tmc3/quantization.h:123-137)"""
    _fields_ = [
        ("num_layers", C.c_int32),
        ("layers", (C.c_int32 * 2) * MAX_QP_LAYERS),
        ("max_qp", C.c_int32),
        ("fixed_point_qp_offset", C.c_int32),
        ("num_ac_coeff_qp_layers", C.c_int32),
        ("ac_coeff_qps", ((C.c_int32 * 2) * 7) * MAX_AC_QP_LAYERS),
    ]


class Predictor(C.Structure):
    _fields_ = [
        ("neighbor_count", C.c_uint32),
        ("predictor_index", C.c_uint32 * 3),
        ("weight", C.c_uint32 * 3),
    ]


MAX_LODS = 32


class LodParams(C.Structure):
    """pccb200_lod_params <- LoD fields of pcc::AttributeParameterSet (tmc3/hls.h:795-857)"""
    _fields_ = [
        ("num_detail_levels", C.c_int32),
        ("lod_decimation_type", C.c_int32),
        ("lod_sampling_period", C.c_int32 * MAX_LODS),
        ("dist2", C.c_int32),
        ("num_pred_nearest_neighbours", C.c_int32),
        ("inter_lod_search_range", C.c_int32),
        ("intra_lod_search_range", C.c_int32),
        ("intra_lod_prediction_skip_layers", C.c_int32),
        ("prediction_with_distribution", C.c_int32),
        ("lod_neigh_bias", C.c_int32 * 3),
        ("pred_weight_blending", C.c_int32),
    ]


PREDICTOR_DTYPE = np.dtype([("neighbor_count", "<u4"), ("predictor_index", "<u4", 3),
                            ("weight", "<u4", 3)])

EXPORTS = [
    "pccb200_abi_version", "pccb200_attr_lift_decode", "pccb200_attr_lift_decode_lod",
    "pccb200_attr_lift_decode_slices", "pccb200_attr_lift_decode_slices_dev",
    "pccb200_attr_lift_encode", "pccb200_attr_lift_encode_lod",
    "pccb200_attr_lift_encode_slices", "pccb200_attr_lift_encode_slices_dev",
    "pccb200_attr_raht_decode",
    "pccb200_attr_raht_decode_multi", "pccb200_attr_raht_decode_multi_batch",
    "pccb200_attr_raht_decode_multi_batch_dev", "pccb200_attr_raht_decode_multi_dev",
    "pccb200_attr_raht_decode_slices_dev", "pccb200_attr_raht_encode",
    "pccb200_attr_raht_encode_multi", "pccb200_attr_raht_encode_multi_batch",
    "pccb200_attr_raht_encode_multi_batch_dev", "pccb200_attr_raht_encode_multi_dev",
    "pccb200_attr_raht_encode_slices", "pccb200_attr_raht_encode_slices_dev",
    "pccb200_attr_raht_encode_symbols", "pccb200_attr_spherical_positions",
    "pccb200_coeff_symbols", "pccb200_estimate_dist2", "pccb200_kernel_launch_count",
    "pccb200_last_error", "pccb200_lift_dequantize", "pccb200_lift_forward",
    "pccb200_lift_inverse", "pccb200_lift_quantize", "pccb200_lod_build", "pccb200_lod_create",
    "pccb200_lod_destroy", "pccb200_lod_info", "pccb200_lod_reusable", "pccb200_morton_sort",
    "pccb200_offset_and_scale", "pccb200_profile_enable", "pccb200_profile_read",
    "pccb200_profile_reset", "pccb200_quant_weights", "pccb200_quant_weights_fixed",
    "pccb200_quant_weights_scalable", "pccb200_raht_forward", "pccb200_raht_inverse",
    "pccb200_raht_params_default", "pccb200_raht_set_prediction_weights", "pccb200_recolour",
    "pccb200_recolour_params_default", "pccb200_set_device",
    "pccb200_time_begin", "pccb200_time_end", "pccb200_xyz_to_rpl",
]
NUM_PHASES = 8
PHASE_NAMES = ["sort", "tree_build", "block_transform", "tail", "gather_scatter", "lifting",
               "block_geometry", "block_schedule"]


class PccB200Error(RuntimeError):
    pass


_lib = None


def lib():
    """Load the CUDA library; raises loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PccB200Error(
                f"{LIB_PATH} not found: build it with `make -C mpeg-pcc-tmc13_b200` "
                "(or __graft_entry__.build()); there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        l.pccb200_last_error.restype = C.c_char_p
        l.pccb200_kernel_launch_count.restype = C.c_uint64
        _lib = l
    return _lib


def _check(rc):
    if rc != 0:
        raise PccB200Error(f"pccb200 status {rc}: {lib().pccb200_last_error().decode()}")


def _p(a, t):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):  # torch CPU tensor (e.g. pinned)
        return C.cast(a.data_ptr(), C.POINTER(t))
    return a.ctypes.data_as(C.POINTER(t))


def default_params():
    p = RahtParams()
    lib().pccb200_raht_params_default(C.byref(p))
    return p


def kernel_launch_count():
    return int(lib().pccb200_kernel_launch_count())


def set_device(i):
    _check(lib().pccb200_set_device(C.c_int(i)))


def morton_sort(xyz):
    xyz = np.ascontiguousarray(xyz, dtype=np.int32)
    n = xyz.shape[0]
    keys = np.empty(n, dtype=np.int64)
    order = np.empty(n, dtype=np.int32)
    _check(lib().pccb200_morton_sort(_p(xyz, C.c_int32), C.c_int32(n), _p(keys, C.c_int64),
                                     _p(order, C.c_int32)))
    return keys, order


def raht_forward(params, qpset, morton, attrs, qpoffs=None):
    """-> (reconstructed attrs [N,A], coefficients [A,N])"""
    attrs = np.ascontiguousarray(attrs, dtype=np.int32).copy()
    n, a = attrs.shape
    morton = np.ascontiguousarray(morton, dtype=np.int64)
    coeffs = np.empty((a, n), dtype=np.int32)
    if qpoffs is not None:
        qpoffs = np.ascontiguousarray(qpoffs, dtype=np.int32)
    _check(lib().pccb200_raht_forward(C.byref(params), C.byref(qpset), _p(qpoffs, C.c_int32),
                                      _p(morton, C.c_int64), _p(attrs, C.c_int32),
                                      C.c_int32(a), C.c_int32(n), _p(coeffs, C.c_int32)))
    return attrs, coeffs


def raht_inverse(params, qpset, morton, coeffs, qpoffs=None):
    coeffs = np.ascontiguousarray(coeffs, dtype=np.int32)
    a, n = coeffs.shape
    morton = np.ascontiguousarray(morton, dtype=np.int64)
    attrs = np.empty((n, a), dtype=np.int32)
    if qpoffs is not None:
        qpoffs = np.ascontiguousarray(qpoffs, dtype=np.int32)
    _check(lib().pccb200_raht_inverse(C.byref(params), C.byref(qpset), _p(qpoffs, C.c_int32),
                                      _p(morton, C.c_int64), _p(attrs, C.c_int32),
                                      C.c_int32(a), C.c_int32(n), _p(coeffs, C.c_int32)))
    return attrs


def attr_raht_encode_into(params, qpset, xyz, attrs_inout, coeffs_out, bitdepth=8,
                          qpoffs=None, slice_offsets=None):
    """Zero-copy form used by bench.py: arrays may be numpy or (pinned) torch
    CPU tensors; attrs_inout [N,A] int32 is overwritten with the clipped
    reconstruction, coeffs_out [A,N] int32 receives the coefficients."""
    n, a = attrs_inout.shape
    if slice_offsets is None:
        _check(lib().pccb200_attr_raht_encode(
            C.byref(params), C.byref(qpset), _p(qpoffs, C.c_int32), _p(xyz, C.c_int32),
            _p(attrs_inout, C.c_int32), C.c_int32(a), C.c_int32(n), C.c_int32(bitdepth),
            _p(coeffs_out, C.c_int32)))
    else:
        so = np.ascontiguousarray(slice_offsets, dtype=np.int64)
        _check(lib().pccb200_attr_raht_encode_slices(
            C.byref(params), C.byref(qpset), _p(qpoffs, C.c_int32), _p(xyz, C.c_int32),
            _p(attrs_inout, C.c_int32), C.c_int32(a), C.c_int32(bitdepth),
            _p(so, C.c_int64), C.c_int32(len(so) - 1), _p(coeffs_out, C.c_int32)))


def attr_raht_encode(params, qpset, xyz, attrs, bitdepth=8, qpoffs=None, slice_offsets=None):
    xyz = np.ascontiguousarray(xyz, dtype=np.int32)
    attrs = np.ascontiguousarray(attrs, dtype=np.int32).copy()
    n, a = attrs.shape
    coeffs = np.empty((a, n), dtype=np.int32)
    if qpoffs is not None:
        qpoffs = np.ascontiguousarray(qpoffs, dtype=np.int32)
    attr_raht_encode_into(params, qpset, xyz, attrs, coeffs, bitdepth, qpoffs, slice_offsets)
    return attrs, coeffs


def attr_raht_decode(params, qpset, xyz, coeffs, bitdepth=8, qpoffs=None):
    xyz = np.ascontiguousarray(xyz, dtype=np.int32)
    coeffs = np.ascontiguousarray(coeffs, dtype=np.int32)
    a, n = coeffs.shape
    attrs = np.empty((n, a), dtype=np.int32)
    if qpoffs is not None:
        qpoffs = np.ascontiguousarray(qpoffs, dtype=np.int32)
    _check(lib().pccb200_attr_raht_decode(
        C.byref(params), C.byref(qpset), _p(qpoffs, C.c_int32), _p(xyz, C.c_int32),
        _p(attrs, C.c_int32), C.c_int32(a), C.c_int32(n), C.c_int32(bitdepth),
        _p(coeffs, C.c_int32)))
    return attrs


def _multi_args(qpsets, arrays_attrs, arrays_coef, bitdepths):
    k = len(qpsets)
    QP = C.POINTER(QpSet) * k
    IP = C.POINTER(C.c_int32) * k
    qp = QP(*[C.pointer(q) for q in qpsets])
    at = IP(*[_p(a, C.c_int32) for a in arrays_attrs])
    co = IP(*[_p(c, C.c_int32) for c in arrays_coef])
    na = (C.c_int32 * k)(*[int(a.shape[1]) for a in arrays_attrs])
    bd = (C.c_int32 * k)(*bitdepths)
    return qp, at, co, na, bd


def attr_raht_encode_multi_into(params, qpsets, xyz, attrs_inout, coeffs_out, bitdepths=None):
    """Several attributes of one slice in one pass (zero-copy form).
    attrs_inout[s]: [N, A_s] int32 (overwritten with the reconstruction),
    coeffs_out[s]: [A_s, N] int32."""
    k = len(qpsets)
    bitdepths = bitdepths or [8] * k
    n = attrs_inout[0].shape[0]
    qp, at, co, na, bd = _multi_args(qpsets, attrs_inout, coeffs_out, bitdepths)
    _check(lib().pccb200_attr_raht_encode_multi(
        C.byref(params), C.c_int32(k), qp, _p(xyz, C.c_int32), at, na, bd, C.c_int32(n), co))


def attr_raht_encode_multi(params, qpsets, xyz, attrs, bitdepths=None):
    xyz = np.ascontiguousarray(xyz, dtype=np.int32)
    a = [np.ascontiguousarray(x, dtype=np.int32).copy() for x in attrs]
    c = [np.empty((x.shape[1], x.shape[0]), dtype=np.int32) for x in a]
    attr_raht_encode_multi_into(params, qpsets, xyz, a, c, bitdepths)
    return a, c


def attr_raht_decode_multi(params, qpsets, xyz, coeffs, bitdepths=None):
    xyz = np.ascontiguousarray(xyz, dtype=np.int32)
    k = len(qpsets)
    bitdepths = bitdepths or [8] * k
    c = [np.ascontiguousarray(x, dtype=np.int32) for x in coeffs]
    a = [np.empty((x.shape[1], x.shape[0]), dtype=np.int32) for x in c]
    n = a[0].shape[0]
    qp, at, co, na, bd = _multi_args(qpsets, a, c, bitdepths)
    _check(lib().pccb200_attr_raht_decode_multi(
        C.byref(params), C.c_int32(k), qp, _p(xyz, C.c_int32), at, na, bd, C.c_int32(n), co))
    return a


def attr_raht_encode_multi_dev(params, qpsets, d_xyz, d_attrs, d_coefs, n, num_attrs, bitdepths=None):
    """device pointers (ints) for xyz, each attribute array and each coefficient array"""
    k = len(qpsets)
    bitdepths = bitdepths or [8] * k
    QP = C.POINTER(QpSet) * k
    VP = C.c_void_p * k
    qp = QP(*[C.pointer(q) for q in qpsets])
    at = VP(*d_attrs)
    co = VP(*d_coefs)
    na = (C.c_int32 * k)(*num_attrs)
    bd = (C.c_int32 * k)(*bitdepths)
    _check(lib().pccb200_attr_raht_encode_multi_dev(
        C.byref(params), C.c_int32(k), qp, C.c_void_p(d_xyz), at, na, bd, C.c_int32(n), co))


def _batch_args(qpsets, units_attrs, units_coefs, bitdepths):
    k = len(qpsets)
    m = len(units_attrs)
    QP = C.POINTER(QpSet) * k
    VP = C.c_void_p * (m * k)
    qp = QP(*[C.pointer(q) for q in qpsets])

    def ptr(x):
        return x if isinstance(x, int) else x.data_ptr() if hasattr(x, "data_ptr") else x.ctypes.data

    at = VP(*[ptr(a) for u in units_attrs for a in u])
    co = VP(*[ptr(c) for u in units_coefs for c in u])
    bd = (C.c_int32 * k)(*bitdepths)
    return qp, at, co, bd


def attr_raht_encode_multi_batch_into(params, qpsets, xyzs, attrs_inout, coeffs_out, bitdepths=None):
    """Many coding units (slices / frames) in one call, zero-copy form.
    xyzs[u]: [N_u, 3] int32; attrs_inout[u][s]: [N_u, A_s] int32 (overwritten
    with the reconstruction); coeffs_out[u][s]: [A_s, N_u] int32."""
    k = len(qpsets)
    m = len(xyzs)
    bitdepths = bitdepths or [8] * k
    qp, at, co, bd = _batch_args(qpsets, attrs_inout, coeffs_out, bitdepths)
    na = (C.c_int32 * k)(*[int(a.shape[1]) for a in attrs_inout[0]])
    xp = (C.c_void_p * m)(*[x.data_ptr() if hasattr(x, "data_ptr") else x.ctypes.data for x in xyzs])
    ns = (C.c_int32 * m)(*[int(x.shape[0]) for x in xyzs])
    _check(lib().pccb200_attr_raht_encode_multi_batch(
        C.byref(params), C.c_int32(k), qp, C.c_int32(m), xp, at, na, bd, ns, co))


def attr_raht_encode_multi_batch(params, qpsets, xyzs, attrs, bitdepths=None):
    """-> (recs[u][s], coefs[u][s])"""
    xyzs = [np.ascontiguousarray(x, dtype=np.int32) for x in xyzs]
    a = [[np.ascontiguousarray(x, dtype=np.int32).copy() for x in u] for u in attrs]
    c = [[np.empty((x.shape[1], x.shape[0]), dtype=np.int32) for x in u] for u in a]
    attr_raht_encode_multi_batch_into(params, qpsets, xyzs, a, c, bitdepths)
    return a, c


def attr_raht_decode_multi_batch(params, qpsets, xyzs, coeffs, bitdepths=None):
    """-> recs[u][s]"""
    k = len(qpsets)
    m = len(xyzs)
    bitdepths = bitdepths or [8] * k
    xyzs = [np.ascontiguousarray(x, dtype=np.int32) for x in xyzs]
    c = [[np.ascontiguousarray(x, dtype=np.int32) for x in u] for u in coeffs]
    a = [[np.empty((x.shape[1], x.shape[0]), dtype=np.int32) for x in u] for u in c]
    qp, at, co, bd = _batch_args(qpsets, a, c, bitdepths)
    na = (C.c_int32 * k)(*[int(x.shape[1]) for x in a[0]])
    xp = (C.c_void_p * m)(*[x.ctypes.data for x in xyzs])
    ns = (C.c_int32 * m)(*[int(x.shape[0]) for x in xyzs])
    _check(lib().pccb200_attr_raht_decode_multi_batch(
        C.byref(params), C.c_int32(k), qp, C.c_int32(m), xp, at, na, bd, ns, co))
    return a


def attr_raht_multi_batch_dev(forward, params, qpsets, d_xyzs, d_attrs, d_coefs, ns, num_attrs,
                              bitdepths=None):
    """device pointers (ints): d_xyzs[u], d_attrs[u][s], d_coefs[u][s]; ns[u] points"""
    k = len(qpsets)
    m = len(d_xyzs)
    bitdepths = bitdepths or [8] * k
    qp, at, co, bd = _batch_args(qpsets, d_attrs, d_coefs, bitdepths)
    na = (C.c_int32 * k)(*num_attrs)
    xp = (C.c_void_p * m)(*d_xyzs)
    nn = (C.c_int32 * m)(*ns)
    fn = (lib().pccb200_attr_raht_encode_multi_batch_dev if forward
          else lib().pccb200_attr_raht_decode_multi_batch_dev)
    _check(fn(C.byref(params), C.c_int32(k), qp, C.c_int32(m), xp, at, na, bd, nn, co))


class RecolourParams(C.Structure):
    _fields_ = [("dist_offset_fwd", C.c_double), ("dist_offset_bwd", C.c_double),
                ("max_geometry_dist2_fwd", C.c_double), ("max_geometry_dist2_bwd", C.c_double),
                ("max_attribute_dist2_fwd", C.c_double), ("max_attribute_dist2_bwd", C.c_double),
                ("search_range", C.c_int32), ("num_neighbours_fwd", C.c_int32),
                ("num_neighbours_bwd", C.c_int32), ("use_dist_weighted_avg_fwd", C.c_int32),
                ("use_dist_weighted_avg_bwd", C.c_int32),
                ("skip_avg_if_identical_source_point_present_fwd", C.c_int32),
                ("skip_avg_if_identical_source_point_present_bwd", C.c_int32),
                ("reserved", C.c_int32)]


def default_recolour_params():
    p = RecolourParams()
    lib().pccb200_recolour_params_default(C.byref(p))
    return p


def recolour(params, source_xyz, source_attrs, target_xyz, scale=1.0, offset=(0, 0, 0), bitdepth=8):
    """attribute transfer source -> target (pccb200_recolour); -> [n_target, A] int32"""
    sx = np.ascontiguousarray(source_xyz, dtype=np.int32)
    sa = np.ascontiguousarray(source_attrs, dtype=np.int32)
    tx = np.ascontiguousarray(target_xyz, dtype=np.int32)
    if sa.ndim == 1:
        sa = sa[:, None]
    a = sa.shape[1]
    out = np.zeros((tx.shape[0], a), dtype=np.int32)
    off = (C.c_int32 * 3)(*[int(v) for v in offset])
    _check(lib().pccb200_recolour(C.byref(params), _p(sx, C.c_int32), _p(sa, C.c_int32), C.c_int32(a),
                                  C.c_int32(sx.shape[0]), C.c_double(scale), off, _p(tx, C.c_int32),
                                  C.c_int32(tx.shape[0]), C.c_int32(bitdepth), _p(out, C.c_int32)))
    return out


def quant_weights(preds, num_points_in_lod):
    preds = np.ascontiguousarray(preds, dtype=PREDICTOR_DTYPE)
    npl = np.ascontiguousarray(num_points_in_lod, dtype=np.uint32)
    n = preds.shape[0]
    qw = np.empty(n, dtype=np.uint64)
    _check(lib().pccb200_quant_weights(C.cast(preds.ctypes.data, C.POINTER(Predictor)),
                                       C.c_int32(n), _p(npl, C.c_uint32), C.c_int32(len(npl)),
                                       _p(qw, C.c_uint64)))
    return qw


def lift(forward, preds, qw, num_points_in_lod, attrs):
    preds = np.ascontiguousarray(preds, dtype=PREDICTOR_DTYPE)
    npl = np.ascontiguousarray(num_points_in_lod, dtype=np.uint32)
    qw = np.ascontiguousarray(qw, dtype=np.uint64)
    attrs = np.ascontiguousarray(attrs, dtype=np.int64).copy()
    if attrs.ndim == 1:
        attrs = attrs[:, None]
    n, a = attrs.shape
    fn = lib().pccb200_lift_forward if forward else lib().pccb200_lift_inverse
    _check(fn(C.cast(preds.ctypes.data, C.POINTER(Predictor)), _p(qw, C.c_uint64), C.c_int32(n),
              _p(npl, C.c_uint32), C.c_int32(len(npl)), _p(attrs, C.c_int64), C.c_int32(a)))
    return attrs


# ---- device-resident entry points (pointers are raw device addresses) ------

def time_begin():
    """Start of a device-timed region spanning every lane (CUDA events)."""
    _check(lib().pccb200_time_begin())


def time_end():
    """-> elapsed device milliseconds since time_begin()."""
    ms = C.c_double(0)
    _check(lib().pccb200_time_end(C.byref(ms)))
    return float(ms.value)


def attr_raht_encode_dev(params, qpset, d_xyz, d_attrs_inout, d_coeffs, n, a, bitdepth=8,
                         d_qpoffs=0, slice_offsets=None):
    so = np.ascontiguousarray(slice_offsets if slice_offsets is not None else [0, n],
                              dtype=np.int64)
    _check(lib().pccb200_attr_raht_encode_slices_dev(
        C.byref(params), C.byref(qpset), C.c_void_p(d_qpoffs or None), C.c_void_p(d_xyz),
        C.c_void_p(d_attrs_inout), C.c_int32(a), C.c_int32(bitdepth), _p(so, C.c_int64),
        C.c_int32(len(so) - 1), C.c_void_p(d_coeffs)))


def attr_raht_decode_dev(params, qpset, d_xyz, d_attrs_out, d_coeffs, n, a, bitdepth=8,
                         d_qpoffs=0, slice_offsets=None):
    so = np.ascontiguousarray(slice_offsets if slice_offsets is not None else [0, n],
                              dtype=np.int64)
    _check(lib().pccb200_attr_raht_decode_slices_dev(
        C.byref(params), C.byref(qpset), C.c_void_p(d_qpoffs or None), C.c_void_p(d_xyz),
        C.c_void_p(d_attrs_out), C.c_int32(a), C.c_int32(bitdepth), _p(so, C.c_int64),
        C.c_int32(len(so) - 1), C.c_void_p(d_coeffs)))


def profile_enable(on):
    lib().pccb200_profile_enable(C.c_int(1 if on else 0))


def profile_reset():
    lib().pccb200_profile_reset()


def profile_read():
    ms = (C.c_double * NUM_PHASES)()
    ln = (C.c_uint64 * NUM_PHASES)()
    lib().pccb200_profile_read(ms, ln)
    return {PHASE_NAMES[i]: (float(ms[i]), int(ln[i])) for i in range(NUM_PHASES)}


def lod_build(params, xyz):
    """-> (predictors[N] (PREDICTOR_DTYPE), indexes[N], num_points_in_lod[lods])"""
    xyz = np.ascontiguousarray(xyz, dtype=np.int32)
    n = xyz.shape[0]
    preds = np.zeros(n, dtype=PREDICTOR_DTYPE)
    indexes = np.zeros(n, dtype=np.uint32)
    npl = np.zeros(MAX_LODS, dtype=np.uint32)
    cnt = C.c_int32(0)
    _check(lib().pccb200_lod_build(C.byref(params), _p(xyz, C.c_int32), C.c_int32(n),
                                   C.cast(preds.ctypes.data, C.POINTER(Predictor)),
                                   _p(indexes, C.c_uint32), _p(npl, C.c_uint32), C.byref(cnt)))
    return preds, indexes, npl[:cnt.value].copy()


def attr_lift_encode(lod_params, qpset, xyz, attrs, lcp_enabled=0, bitdepth=8, qpoffs=None):
    """-> (values [N,A] coding order, reconstruction [N,A] point order, lcp coefficients)"""
    xyz = np.ascontiguousarray(xyz, dtype=np.int32)
    attrs = np.ascontiguousarray(attrs, dtype=np.int32).copy()
    n, a = attrs.shape
    values = np.zeros((n, a), dtype=np.int32)
    lcp = np.zeros(MAX_LODS, dtype=np.int8)
    if qpoffs is not None:
        qpoffs = np.ascontiguousarray(qpoffs, dtype=np.int32)
    _check(lib().pccb200_attr_lift_encode(
        C.byref(lod_params), C.byref(qpset), C.c_int32(lcp_enabled), _p(qpoffs, C.c_int32),
        _p(xyz, C.c_int32), _p(attrs, C.c_int32), C.c_int32(a), C.c_int32(n), C.c_int32(bitdepth),
        _p(values, C.c_int32), _p(lcp, C.c_int8)))
    return values, attrs, lcp[:lod_params.num_detail_levels].copy()


def attr_lift_slices_dev(forward, lod_params, qpset, lcp_enabled, d_xyz, d_attrs, a, slice_offsets,
                         d_values, lcp, bitdepth=8, d_qpoffs=None):
    """device pointers (ints) for xyz / attrs (coded in place) / values; slice_offsets
    (host, int64, numSlices + 1); lcp: host int8 [numSlices, MAX_LODS] (out when forward)"""
    so = np.ascontiguousarray(slice_offsets, dtype=np.int64)
    ns = len(so) - 1
    fn = (lib().pccb200_attr_lift_encode_slices_dev if forward
          else lib().pccb200_attr_lift_decode_slices_dev)
    _check(fn(C.byref(lod_params), C.byref(qpset), C.c_int32(lcp_enabled),
              C.c_void_p(d_qpoffs) if d_qpoffs else None, C.c_void_p(d_xyz), C.c_void_p(d_attrs),
              C.c_int32(a), C.c_int32(bitdepth), _p(so, C.c_int64), C.c_int32(ns),
              C.c_void_p(d_values), _p(lcp, C.c_int8)))


def attr_lift_decode(lod_params, qpset, xyz, values, lcp=None, bitdepth=8, qpoffs=None):
    xyz = np.ascontiguousarray(xyz, dtype=np.int32)
    values = np.ascontiguousarray(values, dtype=np.int32)
    n, a = values.shape
    attrs = np.zeros((n, a), dtype=np.int32)
    l2 = None
    if lcp is not None:
        l2 = np.zeros(MAX_LODS, dtype=np.int8)
        l2[:len(lcp)] = lcp
    _check(lib().pccb200_attr_lift_decode(
        C.byref(lod_params), C.byref(qpset), C.c_int32(1 if lcp is not None else 0),
        _p(qpoffs, C.c_int32), _p(xyz, C.c_int32), _p(attrs, C.c_int32), C.c_int32(a), C.c_int32(n),
        C.c_int32(bitdepth), _p(values, C.c_int32), _p(l2, C.c_int8)))
    return attrs


def _i3(v):
    return (C.c_int32 * 3)(*[int(x) for x in v])


def xyz_to_rpl(laser_origin, laser_theta, xyz):
    """convertXyzToRpl -> (rpl [N,3], bbox (min[3], max[3]))"""
    xyz = np.ascontiguousarray(xyz, dtype=np.int32)
    theta = np.ascontiguousarray(laser_theta, dtype=np.int32)
    out = np.zeros_like(xyz)
    bbox = np.zeros(6, dtype=np.int32)
    _check(lib().pccb200_xyz_to_rpl(_i3(laser_origin), _p(theta, C.c_int32), C.c_int32(theta.size),
                                    _p(xyz, C.c_int32), C.c_int64(xyz.shape[0]),
                                    _p(out, C.c_int32), _p(bbox, C.c_int32)))
    return out, (bbox[:3].copy(), bbox[3:].copy())


def offset_and_scale(min_pos, axis_weight, pos):
    """offsetAndScale -> new positions [N,3]"""
    pos = np.ascontiguousarray(pos, dtype=np.int32).copy()
    _check(lib().pccb200_offset_and_scale(_i3(min_pos), _i3(axis_weight), _p(pos, C.c_int32),
                                          C.c_int64(pos.shape[0])))
    return pos


def attr_spherical_positions(laser_origin, laser_theta, axis_weight, xyz, min_pos=None):
    """conversion + offsetAndScale in one call -> (positions [N,3], bbox of the conversion)"""
    xyz = np.ascontiguousarray(xyz, dtype=np.int32)
    theta = np.ascontiguousarray(laser_theta, dtype=np.int32)
    out = np.zeros_like(xyz)
    bbox = np.zeros(6, dtype=np.int32)
    _check(lib().pccb200_attr_spherical_positions(
        _i3(laser_origin), _p(theta, C.c_int32), C.c_int32(theta.size), _i3(axis_weight),
        None if min_pos is None else _i3(min_pos), _p(xyz, C.c_int32), C.c_int64(xyz.shape[0]),
        _p(out, C.c_int32), _p(bbox, C.c_int32)))
    return out, (bbox[:3].copy(), bbox[3:].copy())


def coeff_symbols(coeffs):
    """planar coefficients [A, N] -> (zero_runs[S], values[S, A], ctx[S] or None, tail_run)"""
    coeffs = np.ascontiguousarray(coeffs, dtype=np.int32)
    a, n = coeffs.shape
    runs = np.zeros(n, dtype=np.int32)
    values = np.zeros((n, a), dtype=np.int32)
    ctx = np.zeros(n, dtype=np.uint8)
    cnt, tail = C.c_int32(0), C.c_int32(0)
    _check(lib().pccb200_coeff_symbols(_p(coeffs, C.c_int32), C.c_int32(a), C.c_int32(n),
                                       _p(runs, C.c_int32), _p(values, C.c_int32),
                                       _p(ctx, C.c_uint8), C.byref(cnt), C.byref(tail)))
    s = cnt.value
    return runs[:s].copy(), values[:s].copy(), (ctx[:s].copy() if a == 3 else None), tail.value


def attr_raht_encode_symbols(params, qpset, xyz, attrs, bitdepth=8, qpoffs=None):
    """-> (reconstruction [N, A], zero_runs, values, ctx, tail_run): the attribute
    encoder call handing over the entropy coder's symbol stream"""
    xyz = np.ascontiguousarray(xyz, dtype=np.int32)
    attrs = np.ascontiguousarray(attrs, dtype=np.int32).copy()
    n, a = attrs.shape
    runs = np.zeros(n, dtype=np.int32)
    values = np.zeros((n, a), dtype=np.int32)
    ctx = np.zeros(n, dtype=np.uint8)
    cnt, tail = C.c_int32(0), C.c_int32(0)
    if qpoffs is not None:
        qpoffs = np.ascontiguousarray(qpoffs, dtype=np.int32)
    _check(lib().pccb200_attr_raht_encode_symbols(
        C.byref(params), C.byref(qpset), _p(qpoffs, C.c_int32), _p(xyz, C.c_int32),
        _p(attrs, C.c_int32), C.c_int32(a), C.c_int32(n), C.c_int32(bitdepth),
        _p(runs, C.c_int32), _p(values, C.c_int32), _p(ctx, C.c_uint8), C.byref(cnt),
        C.byref(tail)))
    s = cnt.value
    return attrs, runs[:s].copy(), values[:s].copy(), (ctx[:s].copy() if a == 3 else None), tail.value


def estimate_dist2(xyz, sampling_period=100, search_range=128, percentile=0.85):
    """estimateDist2 -> shift bits"""
    xyz = np.ascontiguousarray(xyz, dtype=np.int32)
    out = C.c_int32(0)
    _check(lib().pccb200_estimate_dist2(_p(xyz, C.c_int32), C.c_int32(xyz.shape[0]),
                                        C.c_int32(sampling_period), C.c_int32(search_range),
                                        C.c_float(percentile), C.byref(out)))
    return out.value


def quant_weights_fixed(preds, num_points_in_lod, neigh_weight):
    """computeQuantizationWeights (predicting transform) -> qw[N]"""
    preds = np.ascontiguousarray(preds)
    npl = np.ascontiguousarray(num_points_in_lod, dtype=np.uint32)
    n = preds.shape[0]
    qw = np.zeros(n, dtype=np.uint64)
    _check(lib().pccb200_quant_weights_fixed(
        C.cast(preds.ctypes.data, C.POINTER(Predictor)), C.c_int32(n), _p(npl, C.c_uint32),
        C.c_int32(npl.size), _i3(neigh_weight), _p(qw, C.c_uint64)))
    return qw


def quant_weights_scalable(num_points_in_lod, num_points, min_geom_node_size_log2):
    """computeQuantizationWeightsScalable -> qw[N], N = num_points_in_lod[-1]"""
    npl = np.ascontiguousarray(num_points_in_lod, dtype=np.uint32)
    n = int(npl[-1])
    qw = np.zeros(n, dtype=np.uint64)
    _check(lib().pccb200_quant_weights_scalable(
        _p(npl, C.c_uint32), C.c_int32(npl.size), C.c_int64(num_points),
        C.c_int32(min_geom_node_size_log2), C.c_int32(n), _p(qw, C.c_uint64)))
    return qw

