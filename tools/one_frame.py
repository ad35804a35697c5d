"""Developer tool (for ncu captures under gpurun): N fused encodes (colour +
reflectance in one pass) of the bench frame.  one_frame.py [n_points] [calls] [smooth|textured] [enc|dec]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200"))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pcc_attr_b200 as pb  # noqa: E402

bench.N_POINTS = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
textured = (sys.argv[3] if len(sys.argv) > 3 else "textured") == "textured"
decode = (sys.argv[4] if len(sys.argv) > 4 else "enc") == "dec"
xyz, rgb, refl = bench.make_frame(2, textured=textured)
p, q = bench.make_pods(pb)
recs, coefs = pb.attr_raht_encode_multi(p, [q, q], xyz, [rgb, refl])
for _ in range(calls - 1):
    if decode:
        pb.attr_raht_decode_multi(p, [q, q], xyz, coefs)
    else:
        pb.attr_raht_encode_multi(p, [q, q], xyz, [rgb, refl])
