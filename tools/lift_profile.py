"""Developer tool: timing of the lifting path (LoD build, lifting coder) on one GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pcc_attr_b200 as pb
from pcc_testlib import *
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
for name, (xyz, attrs) in (("shell", cloud_shell(n, bits=11, seed=3)), ("lidar", cloud_lidar(n, seed=2))):
    for dec in (0, 1, 2):
        lp = make_lod_params(levels=12, decimation=dec)
        glp = pb.LodParams.from_buffer_copy(bytes(lp))
        qs = make_qpset(qp=34, chroma_offset=-2, fixed_point_qp_offset=24)
        gq = pb.QpSet.from_buffer_copy(bytes(qs))
        pb.lod_build(glp, xyz)
        pb.profile_reset(); pb.profile_enable(True)
        t0 = time.perf_counter(); p, i, npl = pb.lod_build(glp, xyz); t1 = time.perf_counter()
        pb.profile_enable(False); pr = pb.profile_read()
        pb.attr_lift_encode(glp, gq, xyz, attrs, lcp_enabled=1)
        t2 = time.perf_counter(); v, r, l = pb.attr_lift_encode(glp, gq, xyz, attrs, lcp_enabled=1); t3 = time.perf_counter()
        line = f"{name} n={xyz.shape[0]} dec{dec}: lod_build {1e3*(t1-t0):8.1f} ms ({xyz.shape[0]/(t1-t0)/1e6:6.2f} Mpts/s)  lift coder {1e3*(t3-t2):8.1f} ms ({xyz.shape[0]/(t3-t2)/1e6:6.2f} Mpts/s) lods={len(npl)}"
        if liftref_available() and dec == 0:
            t4 = time.perf_counter(); ref_lift_encode(lp, qs, 1, xyz, attrs); t5 = time.perf_counter()
            line += f"  | reference lifting encoder (1 core, incl. entropy coding) {1e3*(t5-t4):8.1f} ms"
        print(line, flush=True)
        print("      kernels: sort %.2f  subsample %.2f  knn %.2f  finalize %.2f ms" % (pr["sort"][0], pr["tree_build"][0], pr["block_transform"][0], pr["tail"][0]), flush=True)
