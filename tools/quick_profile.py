"""Developer tool: per-phase timing of one attribute call in several
configurations (run under gpurun).

  quick_profile.py [n] [lidar|shell] [texture amplitude RGB] [texture amplitude refl]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200"))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import pcc_attr_b200 as pb  # noqa: E402
from pcc_attr_b200.synth import texture  # noqa: E402
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
cloud = sys.argv[2] if len(sys.argv) > 2 else "lidar"
tex_rgb = int(sys.argv[3]) if len(sys.argv) > 3 else 0
tex_refl = int(sys.argv[4]) if len(sys.argv) > 4 else tex_rgb
bench.N_POINTS = n
if cloud == "lidar":
    xyz, rgb, refl = bench.make_frame(2, textured=False)
else:
    from pcc_attr_b200.synth import cloud_shell
    xyz, rgb = cloud_shell(n, bits=11, seed=3)
    refl = rgb[:, :1].copy()
if tex_rgb:
    rgb = texture(rgb, tex_rgb, 77)
if tex_refl:
    refl = texture(refl, tex_refl, 78)
p, q = bench.make_pods(pb)
print(f"# env: " + " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("PCCB200_")),
      f"| texture rgb +-{tex_rgb} refl +-{tex_refl}")


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
    nz = int(np.count_nonzero(np.abs(coef).sum(axis=0) if coef.ndim == 2 else coef))
    print(f"{label:24s} A={a} wall {1e3*(t1-t0):7.1f} ms  block {pr['block_transform'][0]:7.2f} ms "
          f"({pr['block_transform'][1]:3d}) geom {pr['block_geometry'][0]:5.2f} ({pr['block_geometry'][1]:3d}) "
          f"sched {pr['block_schedule'][0]:5.2f} ({pr['block_schedule'][1]:3d}) sort {pr['sort'][0]:.2f} "
          f"tree {pr['tree_build'][0]:.2f} tail {pr['tail'][0]:.2f} gather {pr['gather_scatter'][0]:.2f} nonzero {nz}",
          flush=True)


def run_multi(label, decode=False):
    recs, coefs = pb.attr_raht_encode_multi(p, [q, q], xyz, [rgb, refl])
    fn = (lambda: pb.attr_raht_decode_multi(p, [q, q], xyz, coefs)) if decode else \
        (lambda: pb.attr_raht_encode_multi(p, [q, q], xyz, [rgb, refl]))
    fn()
    pb.profile_reset()
    pb.profile_enable(True)
    t0 = time.perf_counter()
    fn()
    t1 = time.perf_counter()
    pb.profile_enable(False)
    pr = pb.profile_read()
    print(f"{label:24s} A=3+1 wall {1e3*(t1-t0):7.1f} ms  block {pr['block_transform'][0]:7.2f} ms "
          f"({pr['block_transform'][1]:3d}) geom {pr['block_geometry'][0]:5.2f} ({pr['block_geometry'][1]:3d}) "
          f"sched {pr['block_schedule'][0]:5.2f} ({pr['block_schedule'][1]:3d}) sort {pr['sort'][0]:.2f} "
          f"tree {pr['tree_build'][0]:.2f} tail {pr['tail'][0]:.2f} gather {pr['gather_scatter'][0]:.2f}",
          flush=True)


pn = pb.RahtParams.from_buffer_copy(bytes(p)); pn.prediction_enabled = 0
ps = pb.RahtParams.from_buffer_copy(bytes(p)); ps.subnode_prediction_enabled = 0
full = os.environ.get("QP_FULL", "1") != "0"
run_multi("enc multi")
run_multi("dec multi", decode=True)
for a, attrs in ((3, rgb), (1, refl)):
    run("enc default", p, a, attrs)
    run("dec default", p, a, attrs, decode=True)
    if full:
        run("enc nopred (tz only)", pn, a, attrs)
        run("dec nopred (no deps)", pn, a, attrs, decode=True)
        run("enc nosubnode", ps, a, attrs)
