#!/bin/bash
# GPU pass B: parity of the restructured descent, chunk sweep
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 900 python -m pytest tests -m gpu -x -q --timeout=300 --timeout-method=thread > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
tail -4 gpurun_out/b_pytest.log
QP="python tools/quick_profile.py 1000000 lidar"
{
  for tex in "0 0" "16 24" "32 32"; do
    for ch in 1 2 4 8 16; do
      echo "=== texture $tex: chunk $ch"
      PCCB200_BLOCK_CHUNK=$ch QP_FULL=0 timeout -k 10 90 $QP $tex
    done
    echo "=== texture $tex: legacy path chunk 1"
    PCCB200_BLOCK_KERNEL=warp PCCB200_BLOCK_CHUNK=1 QP_FULL=0 timeout -k 10 90 $QP $tex
  done
  echo "=== decoder, Morton order"
  PCCB200_WAVE_ORDER=morton QP_FULL=0 timeout -k 10 90 $QP 0 0
} > gpurun_out/b_profile.log 2>&1
grep -E "===|enc default|dec default" gpurun_out/b_profile.log | cut -c1-150
