#!/bin/bash
# GPU pass P: neighbour tables by actual block count (more units fit), ncu launch list + full capture of the gang kernel
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_batch or multi_attribute or golden or full_size or whole_codec_bitstream" --timeout=200 --timeout-method=thread > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p_pytest.log
tail -4 gpurun_out/p_pytest.log
GANG_SWEEP="128:4,160:5,192:6" timeout -k 10 400 python tools/gang_sweep.py > gpurun_out/p_sweep_textured.log 2>&1
cat gpurun_out/p_sweep_textured.log | tail -5
GANG_SWEEP="128:4,192:6" timeout -k 10 300 python tools/gang_sweep.py 0 0 4 > gpurun_out/p_sweep_smooth.log 2>&1
cat gpurun_out/p_sweep_smooth.log | tail -4
export GANG_NOREF=1 GANG_STEPS=0 GANG_SWEEP="16:4"
timeout -k 10 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv --log-file gpurun_out/p_launches.csv python tools/gang_sweep.py 16 24 2 > gpurun_out/p_ncu_list.log 2>&1
tail -3 gpurun_out/p_ncu_list.log
timeout -k 10 500 ncu --set full --import-source on --clock-control none --kernel-name regex:k_block_warp_gang --launch-skip 90 --launch-count 2 -f -o gpurun_out/p_gang_full python tools/gang_sweep.py 16 24 2 > gpurun_out/p_ncu_full.log 2>&1
tail -3 gpurun_out/p_ncu_full.log
ls -la gpurun_out/p_*
