#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 200 > gpurun_out/c25_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c25_pytest.txt; tail -n 5 gpurun_out/c25_pytest.txt
timeout 300 python tools/ab_matrix.py stock '' > gpurun_out/c25_ab.txt 2>&1; cat gpurun_out/c25_ab.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c25_smoke.txt 2>&1; tail -n 3 gpurun_out/c25_smoke.txt
