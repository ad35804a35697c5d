import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pcc_attr_b200 as pb
from pcc_testlib import *
for xyz, attrs in (cloud_shell(120000, bits=10, seed=7, a=3), cloud_lidar(80000, seed=4, a=3)):
    for dec, lcp, qp in ((0, 1, 34), (2, 1, 22)):
        lp = make_lod_params(levels=10, decimation=dec)
        qs = make_qpset(qp=qp, chroma_offset=-2, fixed_point_qp_offset=24, layers=[(qp, -2), (qp + 2, 0), (qp + 4, 1)])
        ov, orr, ol = oracle_lift_encode(lp, qs, lcp, xyz, attrs)
        gv, gr, gl = pb.attr_lift_encode(pb.LodParams.from_buffer_copy(bytes(lp)), pb.QpSet.from_buffer_copy(bytes(qs)), xyz, attrs, lcp_enabled=lcp)
        p, i, npl = oracle_lod_build(lp, xyz)
        print(dec, qp, "values", np.array_equal(gv, ov), "recon", np.array_equal(gr, orr), "npl", len(npl), "gl", gl, "ol", ol)
