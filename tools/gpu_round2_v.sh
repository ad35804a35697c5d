#!/bin/bash
# GPU pass V: recolouring on the GPU (parity vs oracle, full size), device-pointer lifting entries
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 900 python -m pytest tests/test_recolour.py tests/test_gpu_parity.py -m gpu -x -q -k "recolour or gpu_vs_oracle or gpu_full_size or device_pointers" --timeout=600 --timeout-method=thread > gpurun_out/v_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/v_pytest.log
tail -6 gpurun_out/v_pytest.log
timeout -k 10 300 python - > gpurun_out/v_recolour_time.log 2>&1 <<'PY'
import sys, time, os
sys.path.insert(0, "mpeg-pcc-tmc13_b200"); sys.path.insert(0, "tests")
import numpy as np, pcc_attr_b200 as pb, bench
xyz, rgb, refl = bench.make_frame(2)
half = np.ascontiguousarray(np.unique(np.rint(xyz * 0.5).astype(np.int32), axis=0))
rp = pb.default_recolour_params()
pb.profile_enable(True)
for rep in range(3):
    pb.profile_reset()
    t0 = time.perf_counter(); out = pb.recolour(rp, xyz, rgb, half, 0.5); t = time.perf_counter() - t0
    print(f"recolour 1M lidar -> {half.shape[0]} targets: {1e3*t:.1f} ms wall", {k: round(v[0], 2) for k, v in pb.profile_read().items() if v[0]}, flush=True)
from pcc_attr_b200.synth import cloud_shell
sx, sa = cloud_shell(1000000, bits=11, seed=7)
st = np.ascontiguousarray(np.unique(np.rint(sx * 0.5).astype(np.int32), axis=0))
for rep in range(2):
    pb.profile_reset()
    t0 = time.perf_counter(); out = pb.recolour(rp, sx, sa, st, 0.5); t = time.perf_counter() - t0
    print(f"recolour 1M shell -> {st.shape[0]} targets: {1e3*t:.1f} ms wall", {k: round(v[0], 2) for k, v in pb.profile_read().items() if v[0]}, flush=True)
PY
cat gpurun_out/v_recolour_time.log | tail -6
timeout -k 10 600 python bench.py --workload lift10m --steps 3 --warmup 3 > gpurun_out/v_lift10m.json 2> gpurun_out/v_lift10m.err
echo "lift10m rc=$?"; tail -2 gpurun_out/v_lift10m.err; cut -c1-300 gpurun_out/v_lift10m.json
