"""Developer tool (run under gpurun): throughput of the batch entry point
(pccb200_attr_raht_encode_multi_batch_dev) as a function of the number of
units in flight and the gang size.

  gang_sweep.py [texture amplitude RGB] [texture amplitude refl] [distinct frames]
  env GANG_SWEEP="F:G,F:G,..."   units per step : units per gang (0 = spread over lanes)
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200"))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np  # noqa: E402
import torch  # noqa: E402
import pcc_attr_b200 as pb  # noqa: E402
from pcc_attr_b200.synth import texture  # noqa: E402
import bench  # noqa: E402

tex_rgb = int(sys.argv[1]) if len(sys.argv) > 1 else bench.TEXTURE_RGB
tex_refl = int(sys.argv[2]) if len(sys.argv) > 2 else bench.TEXTURE_REFL
distinct = int(sys.argv[3]) if len(sys.argv) > 3 else 8
def _entry(s):
    parts = s.split(":")
    env = dict(kv.split("=") for kv in parts[2].split(";")) if len(parts) > 2 and parts[2] else {}
    return int(parts[0]), int(parts[1]), env


sweep = [_entry(s) for s in os.environ.get("GANG_SWEEP", "32:1,64:2,128:4,128:8").split(",")]
dev = torch.device("cuda", 0)
p, q = bench.make_pods(pb)
pb.lib()
pb.set_device(0)

base = []
for i in range(distinct):
    xyz, rgb, refl = bench.make_frame(2 + i, textured=False)
    if tex_rgb:
        rgb = texture(rgb, tex_rgb, 2002 + i)
    if tex_refl:
        refl = texture(refl, tex_refl, 3002 + i)
    base.append((torch.from_numpy(xyz).to(dev), torch.from_numpy(rgb).to(dev),
                 torch.from_numpy(refl).to(dev)))
n = base[0][0].shape[0]
print(f"# texture rgb +-{tex_rgb} refl +-{tex_refl}, {distinct} distinct frames of {n} points", flush=True)

# reference result of frame 0 through the single-unit entry (GANG_NOREF=1: skipped,
# GANG_STEPS: timed steps per sweep entry -- for runs under ncu)
NOREF = os.environ.get("GANG_NOREF", "0") == "1"
STEPS = int(os.environ.get("GANG_STEPS", "2"))
ref = {"rgb": base[0][1].clone(), "refl": base[0][2].clone(),
       "crgb": torch.empty((3, n), dtype=torch.int32, device=dev),
       "crefl": torch.empty((1, n), dtype=torch.int32, device=dev)}
t0 = time.perf_counter()
if not NOREF:
    pb.attr_raht_encode_multi_dev(p, [q, q], base[0][0].data_ptr(),
                                  [ref["rgb"].data_ptr(), ref["refl"].data_ptr()],
                                  [ref["crgb"].data_ptr(), ref["crefl"].data_ptr()], n, [3, 1])
torch.cuda.synchronize()
print(f"single-unit call, frame 0: {1e3 * (time.perf_counter() - t0):.1f} ms (first call)", flush=True)

maxF = max(f for f, _, _ in sweep)
units = []
for u in range(maxF):
    x, r, l = base[u % distinct]
    units.append({"xyz": x, "rgb0": r, "refl0": l, "rgb": torch.empty_like(r),
                  "refl": torch.empty_like(l),
                  "crgb": torch.empty((3, n), dtype=torch.int32, device=dev),
                  "crefl": torch.empty((1, n), dtype=torch.int32, device=dev)})


def step(F):
    for d in units[:F]:
        d["rgb"].copy_(d["rgb0"])
        d["refl"].copy_(d["refl0"])
    torch.cuda.synchronize()
    pb.time_begin()
    pb.attr_raht_multi_batch_dev(True, p, [q, q], [d["xyz"].data_ptr() for d in units[:F]],
                                 [[d["rgb"].data_ptr(), d["refl"].data_ptr()] for d in units[:F]],
                                 [[d["crgb"].data_ptr(), d["crefl"].data_ptr()] for d in units[:F]],
                                 [n] * F, [3, 1])
    return pb.time_end()


KNOBS = ("PCCB200_GANG_CTAS", "PCCB200_POLL_NS", "PCCB200_BLOCK_SHARE")
for F, G, env in sweep:
    os.environ["PCCB200_GANG"] = str(G)
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in env.items():
        os.environ["PCCB200_" + k] = v
    try:
        if STEPS > 0:
            step(F)
        ok = NOREF or all(torch.equal(units[0][k], ref[k]) for k in ("rgb", "refl", "crgb", "crefl"))
        last = units[F - 1]
        same = (F - 1) % distinct == 0
        ok_last = NOREF or (not same) or all(torch.equal(last[k], ref[k]) for k in ("rgb", "refl", "crgb", "crefl"))
        ts = [step(F) for _ in range(max(1, STEPS))]
        ms = min(ts)
        free, total = torch.cuda.mem_get_info()
        print(f"units {F:4d} gang {G:3d} {env}: {ms:9.1f} ms/step  {F * n / ms / 1e3:8.1f} Mpoints/s  "
              f"{ms / F:7.2f} ms/unit  bit-exact vs single call: {ok and ok_last}  "
              f"HBM in use {(total - free) / 2**30:.1f} GiB", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"units {F} gang {G}: FAILED {e}", flush=True)
        break
