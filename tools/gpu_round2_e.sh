#!/bin/bash
# GPU pass E: what bounds the textured encoder: window (grid) and poll-interval sweeps
mkdir -p gpurun_out
cd /root/repo
QP="python tools/quick_profile.py 1000000 lidar"
{
  for g in 4 16 64 148 444; do
    for pn in 32 400; do
      echo "=== textured 16 24: grid $g pollNs $pn"
      PCCB200_BLOCK_GRID=$g PCCB200_POLL_NS=$pn QP_FULL=0 timeout -k 10 150 $QP 16 24 | grep -E "enc |dec multi"
    done
  done
  for g in 16 64 148 444; do
    echo "=== smooth: grid $g"
    PCCB200_BLOCK_GRID=$g QP_FULL=0 timeout -k 10 100 $QP 0 0 | grep -E "enc |dec multi"
  done
  echo "=== per-stage times, decoder + encoder, smooth"
  PCCB200_DEBUG=1 QP_FULL=0 timeout -k 10 100 $QP 0 0 2>&1 | grep -E "launches|enc multi|dec multi|enc default|dec default" | cut -c1-400
  echo "=== per-stage times, textured"
  PCCB200_DEBUG=1 QP_FULL=0 timeout -k 10 150 $QP 16 24 2>&1 | grep -E "launches|enc multi|dec multi|enc default|dec default" | cut -c1-400
} > gpurun_out/e_sweep.log 2>&1
cat gpurun_out/e_sweep.log | cut -c1-260
