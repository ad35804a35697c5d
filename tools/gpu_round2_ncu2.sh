#!/bin/bash
# ncu --set full of the distance-subsampling dataflow kernel (first, largest level) and of the recolouring k-NN
mkdir -p gpurun_out
cd /root/repo
timeout -k 10 400 ncu --set full --import-source on --clock-control none --kernel-name regex:k_subsample_cells --launch-count 2 -f -o gpurun_out/n2_subsample python tools/ncu_lift_recolour.py lift > gpurun_out/n2_subsample.log 2>&1
tail -2 gpurun_out/n2_subsample.log
timeout -k 10 400 ncu --set full --import-source on --clock-control none --kernel-name regex:KnnQueryFn --launch-count 2 -f -o gpurun_out/n2_knn python tools/ncu_lift_recolour.py recolour > gpurun_out/n2_knn.log 2>&1
tail -2 gpurun_out/n2_knn.log
ls -la gpurun_out/n2_*
