#!/bin/bash
# GPU pass U: 16-byte cell records in the distance subsampling; pinned host buffers in the lifting workloads; sanitizer
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lod or lift or whole_codec or fuzz" --timeout=300 --timeout-method=thread > gpurun_out/u_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/u_pytest.log
tail -4 gpurun_out/u_pytest.log
for w in predlift3m lift10m; do
  timeout -k 10 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/u_$w.json 2> gpurun_out/u_$w.err
  echo "$w rc=$?"; tail -2 gpurun_out/u_$w.err; cut -c1-260 gpurun_out/u_$w.json
done
TOOLS="memcheck racecheck synccheck" timeout -k 10 900 bash tools/sanitize.sh > gpurun_out/u_sanitize.log 2>&1
tail -30 gpurun_out/u_sanitize.log
