#!/bin/bash
# GPU pass AA: wavefront order for the distance subsampling
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lod or lift or whole_codec or fuzz" --timeout=300 --timeout-method=thread > gpurun_out/aa_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/aa_pytest.log
tail -3 gpurun_out/aa_pytest.log
for wave in 1 0; do
  PCCB200_SUBSAMPLE_WAVE=$wave timeout -k 10 600 python bench.py --workload predlift3m --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/aa_predlift3m_wave$wave.json 2> gpurun_out/aa_predlift3m_wave$wave.err
  echo "predlift3m wave=$wave rc=$?"; python -c "
import json;d=json.loads(open('gpurun_out/aa_predlift3m_wave$wave.json').read().strip().splitlines()[-1]);print(round(d['value'],1), {k:round(v,2) for k,v in d['phase_ms_one_slice_alone'].items() if v})"
done
timeout -k 10 600 python bench.py --workload lift10m --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/aa_lift10m.json 2> gpurun_out/aa_lift10m.err
python -c "
import json;d=json.loads(open('gpurun_out/aa_lift10m.json').read().strip().splitlines()[-1]);print('lift10m', round(d['value'],1), round(d['e2e']['value'],1), {k:round(v,2) for k,v in d['phase_ms_one_slice_alone'].items() if v})"
