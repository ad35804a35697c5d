"""Developer tool: per-phase timing of one attribute call in several
configurations (run under gpurun)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import pcc_attr_b200 as pb  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
cloud = sys.argv[2] if len(sys.argv) > 2 else "lidar"
bench.N_POINTS = n
if cloud == "lidar":
    xyz, rgb, refl = bench.make_frame(2)
else:
    from pcc_attr_b200.synth import cloud_shell
    xyz, rgb = cloud_shell(n, bits=11, seed=3)
    refl = rgb[:, :1].copy()
p, q = bench.make_pods(pb)


def run(label, params, a, attrs, decode=False):
    rec, coef = pb.attr_raht_encode(params, q, xyz, attrs)
    fn = (lambda: pb.attr_raht_decode(params, q, xyz, coef)) if decode else \
        (lambda: pb.attr_raht_encode(params, q, xyz, attrs))
    fn()
    pb.profile_reset()
    pb.profile_enable(True)
    t0 = time.perf_counter()
    fn()
    t1 = time.perf_counter()
    pb.profile_enable(False)
    pr = pb.profile_read()
    print(f"{label:28s} A={a} n={xyz.shape[0]} wall {1e3*(t1-t0):8.1f} ms  block {pr['block_transform'][0]:8.2f} ms "
          f"({pr['block_transform'][1]} launches) sort {pr['sort'][0]:.2f} tree {pr['tree_build'][0]:.2f}")


import copy
pn = pb.RahtParams.from_buffer_copy(bytes(p)); pn.prediction_enabled = 0
ps = pb.RahtParams.from_buffer_copy(bytes(p)); ps.subnode_prediction_enabled = 0
for a, attrs in ((3, rgb), (1, refl)):
    run("enc default", p, a, attrs)
    run("dec default", p, a, attrs, decode=True)
    run("enc nopred (tz only)", pn, a, attrs)
    run("dec nopred (no deps)", pn, a, attrs, decode=True)
    run("enc nosubnode", ps, a, attrs)
