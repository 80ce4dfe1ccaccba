#!/bin/bash
# round-2 GPU call #2: chain kernel bring-up, full parity suite, A/B chain on/off, bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "chain_kernel" --timeout 120 > gpurun_out/c2_chain.txt 2>&1; rc=$?; echo "chain rc=$rc" >> gpurun_out/c2_chain.txt
tail -n 5 gpurun_out/c2_chain.txt
if [ $rc -ne 0 ]; then
  # bring-up diagnostics: smallest case under the sanitizer tools
  timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "chain_kernel and stock and float16 and 3x96x64" --timeout 200 > gpurun_out/c2_chain_memcheck.txt 2>&1
  tail -n 30 gpurun_out/c2_chain_memcheck.txt
fi
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_parity.py::test_chain_kernel_matches_per_layer_kernels > gpurun_out/c2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c2_pytest.txt
tail -n 8 gpurun_out/c2_pytest.txt
timeout 600 python tools/ab_matrix.py stock '' 'chain=0' > gpurun_out/c2_ab.txt 2>&1
timeout 300 python tools/ab_matrix.py pruned '' 'chain=0' >> gpurun_out/c2_ab.txt 2>&1
cat gpurun_out/c2_ab.txt
timeout 600 python bench.py > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
