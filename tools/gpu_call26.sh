#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=$PWD/fastdepth_b200/libfastdepth_b200_bigsmem.so
FD_B200_LIB=$V timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 200 -k "golden or stage_by_stage or baseline_configs or agree_bitwise or chain_kernel" > gpurun_out/c26_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c26_pytest.txt; tail -n 4 gpurun_out/c26_pytest.txt
echo "== main" > gpurun_out/c26_ab.txt
timeout 200 python tools/ab_matrix.py stock '' >> gpurun_out/c26_ab.txt 2>&1
echo "== bigsmem + halfk unroll" >> gpurun_out/c26_ab.txt
FD_B200_LIB=$V timeout 200 python tools/ab_matrix.py stock '' >> gpurun_out/c26_ab.txt 2>&1
echo "== main" >> gpurun_out/c26_ab.txt
timeout 200 python tools/ab_matrix.py stock '' >> gpurun_out/c26_ab.txt 2>&1
echo "== bigsmem + halfk unroll" >> gpurun_out/c26_ab.txt
FD_B200_LIB=$V timeout 200 python tools/ab_matrix.py stock '' >> gpurun_out/c26_ab.txt 2>&1
echo "== pruned main / variant" >> gpurun_out/c26_ab.txt
timeout 200 python tools/ab_matrix.py pruned '' >> gpurun_out/c26_ab.txt 2>&1
FD_B200_LIB=$V timeout 200 python tools/ab_matrix.py pruned '' >> gpurun_out/c26_ab.txt 2>&1
cat gpurun_out/c26_ab.txt
