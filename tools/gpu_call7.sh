#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "chain_kernel" --timeout 120 > gpurun_out/c7_chain.txt 2>&1; echo "chain rc=$?" >> gpurun_out/c7_chain.txt
tail -n 4 gpurun_out/c7_chain.txt
timeout 200 python tools/trace_chain.py > gpurun_out/c7_trace_chain.txt 2>&1; cat gpurun_out/c7_trace_chain.txt
echo "== main (FHFMA)" > gpurun_out/c7_ab.txt
timeout 300 python tools/ab_matrix.py stock '' >> gpurun_out/c7_ab.txt 2>&1
echo "== variant FFMA2" >> gpurun_out/c7_ab.txt
FD_B200_LIB=$PWD/fastdepth_b200/libfastdepth_b200_chainffma2.so timeout 300 python tools/ab_matrix.py stock '' >> gpurun_out/c7_ab.txt 2>&1
FD_B200_LIB=$PWD/fastdepth_b200/libfastdepth_b200_chainffma2.so timeout 200 python tools/trace_chain.py > gpurun_out/c7_trace_chain_ffma2.txt 2>&1
timeout 300 python tools/ab_matrix.py pruned '' >> gpurun_out/c7_ab.txt 2>&1
cat gpurun_out/c7_ab.txt
