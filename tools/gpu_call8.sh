#!/bin/bash
# cluster-mode bring-up of the block kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for cl in 4 2; do
  FD_TC_CLUSTER=$cl timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_stage_by_stage and float16 and 1-1" --timeout 120 > gpurun_out/c8_small_cl$cl.txt 2>&1; echo "small cl$cl rc=$?" >> gpurun_out/c8_small_cl$cl.txt
  tail -n 5 gpurun_out/c8_small_cl$cl.txt
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "agree_bitwise" --timeout 300 > gpurun_out/c8_bitwise.txt 2>&1; echo "bitwise rc=$?" >> gpurun_out/c8_bitwise.txt
tail -n 12 gpurun_out/c8_bitwise.txt
timeout 900 python tools/ab_matrix.py stock '' 'FD_TC_CLUSTER=1' 'FD_TC_CLUSTER=2' 'FD_TC_CLUSTER=4' 'chain=0' 'chain=0,FD_TC_CLUSTER=1' 'chain=0,FD_TC_CLUSTER=2' 'chain=0,FD_TC_CLUSTER=4' > gpurun_out/c8_ab.txt 2>&1
cat gpurun_out/c8_ab.txt
timeout 600 python tools/ab_matrix.py pruned '' 'FD_TC_CLUSTER=1' 'FD_TC_CLUSTER=2' 'FD_TC_CLUSTER=4' > gpurun_out/c8_ab_pruned.txt 2>&1
cat gpurun_out/c8_ab_pruned.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c8_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c8_pytest.txt
tail -n 8 gpurun_out/c8_pytest.txt
