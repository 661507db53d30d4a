#!/bin/bash
# ncu --set full captures: one MHA forward launch, the MN-major pair-GEMM launches of one training step (bounded counts)
set -x
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:mha_fwd_tc64 --launch-skip 3 --launch-count 1 \
    -o gpurun_out/mha_full -f python bench.py --profile-step > gpurun_out/ncu_mha.log 2>&1
echo rc=$?
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:pair_kernel<256, 5, 0, true, true>' --launch-count 3 \
    -o gpurun_out/pair_mnmn_full -f python bench.py --profile-step > gpurun_out/ncu_pair_mn.log 2>&1
echo rc=$?
ls -la gpurun_out/*.ncu-rep
