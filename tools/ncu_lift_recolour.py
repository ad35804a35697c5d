"""Developer tool (run under ncu / gpurun): one lifting encode of a 1M-point
surface slice (LoD build: distance subsampling, neighbour search; lifting;
quantisation) and one recolouring of a 1M-point frame, for a launch list."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import pcc_attr_b200 as pb  # noqa: E402
from pcc_attr_b200.synth import cloud_shell, texture  # noqa: E402
import bench  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "both"
xyz, rgb = cloud_shell(1000000, bits=11, seed=40)
rgb = texture(rgb, 16, 41)
if what in ("lift", "both"):
    lp = pb.LodParams()
    lp.num_detail_levels, lp.lod_decimation_type, lp.dist2 = 12, 0, 0
    lp.num_pred_nearest_neighbours, lp.inter_lod_search_range = 3, 1100000
    lp.intra_lod_search_range, lp.intra_lod_prediction_skip_layers = 0, 13
    lp.prediction_with_distribution, lp.pred_weight_blending = 1, 0
    for i in range(3):
        lp.lod_neigh_bias[i] = 1
    for i in range(32):
        lp.lod_sampling_period[i] = 4
    lq = pb.QpSet()
    lq.num_layers, lq.max_qp, lq.fixed_point_qp_offset = 1, 51, 24
    lq.layers[0][0], lq.layers[0][1] = bench.QP, bench.CHROMA_OFFSET
    pb.attr_lift_encode(lp, lq, xyz, rgb, lcp_enabled=1)
    print("lifting encode done", flush=True)
if what in ("recolour", "both"):
    half = np.ascontiguousarray(np.unique(np.rint(xyz * 0.5).astype(np.int32), axis=0))
    out = pb.recolour(pb.default_recolour_params(), xyz, rgb, half, 0.5)
    print("recolour done", out.shape, flush=True)
