#!/bin/bash
# session-2 call #1: re-establish state (chain parity, trace, A/B chain on/off, full suite, bench)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "chain_kernel" --timeout 120 > gpurun_out/c5_chain.txt 2>&1; echo "chain rc=$?" >> gpurun_out/c5_chain.txt
tail -n 6 gpurun_out/c5_chain.txt
timeout 200 python tools/trace_chain.py > gpurun_out/c5_trace_chain.txt 2>&1; cat gpurun_out/c5_trace_chain.txt
timeout 400 python tools/ab_matrix.py stock '' 'chain=0' > gpurun_out/c5_ab.txt 2>&1
timeout 300 python tools/ab_matrix.py pruned '' 'chain=0' >> gpurun_out/c5_ab.txt 2>&1
cat gpurun_out/c5_ab.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_parity.py::test_chain_kernel_matches_per_layer_kernels > gpurun_out/c5_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c5_pytest.txt
tail -n 8 gpurun_out/c5_pytest.txt
timeout 600 python bench.py > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; tail -c 600 gpurun_out/c5_bench.err
