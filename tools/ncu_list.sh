#!/bin/bash
# developer tool (run under gpurun): ncu launch list (per-launch durations and
# DRAM bytes) of 2 x (RGB encode + reflectance encode) of the bench frame
mkdir -p gpurun_out
cat > /tmp/one_call.py <<'PY'
import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, os.path.join(ROOT, "mpeg-pcc-tmc13_b200")); sys.path.insert(0, ROOT)
import pcc_attr_b200 as pb, bench
bench.N_POINTS = int(sys.argv[1])
xyz, rgb, refl = bench.make_frame(2)
p, q = bench.make_pods(pb)
for _ in range(int(sys.argv[2])):
    pb.attr_raht_encode(p, q, xyz, rgb)
    pb.attr_raht_encode(p, q, xyz, refl)
PY
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv \
  --log-file gpurun_out/${OUT:-launches.csv} python /tmp/one_call.py 1000000 2 > gpurun_out/ncu_list.log 2>&1
tail -1 gpurun_out/ncu_list.log
