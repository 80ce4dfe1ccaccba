#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 60 tools/umma_rate > gpurun_out/c13_umma_rate.txt 2>&1; cat gpurun_out/c13_umma_rate.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/c13_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c13_pytest.txt
tail -n 8 gpurun_out/c13_pytest.txt
timeout 300 python tools/ab_matrix.py stock '' > gpurun_out/c13_ab.txt 2>&1; cat gpurun_out/c13_ab.txt
timeout 600 python bench.py > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err; tail -c 300 gpurun_out/c13_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c13_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['bound'], d['roofline']['frac'], d['roofline']['kernel'][:40])"
