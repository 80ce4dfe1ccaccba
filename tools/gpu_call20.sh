#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -x --timeout 200 > gpurun_out/c20_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c20_pytest.txt; tail -n 6 gpurun_out/c20_pytest.txt
timeout 400 python tools/ab_matrix.py stock 'FD_STEM_NO_TMA_EPI=1' '' 'FD_STEM_NO_TMA_EPI=1' '' > gpurun_out/c20_ab.txt 2>&1; cat gpurun_out/c20_ab.txt
timeout 200 python tools/ab_matrix.py pruned 'FD_STEM_NO_TMA_EPI=1' '' > gpurun_out/c20_ab_pruned.txt 2>&1; cat gpurun_out/c20_ab_pruned.txt
