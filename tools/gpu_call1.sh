#!/bin/bash
# round-2 GPU call #1: parity suite, 2-CTA smoke, A/B matrix, stage traces, bench line, sanitizer quick pass
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 120 tools/umma2_smoke > gpurun_out/c1_umma2.txt 2>&1; echo "umma2 rc=$?" >> gpurun_out/c1_umma2.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/c1_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c1_pytest.txt
timeout 600 python tools/ab_matrix.py > gpurun_out/c1_ab.txt 2>&1
timeout 300 python tools/trace_stage.py 1 2 3 4 5 6 7 12 13 14 15 16 17 18 > gpurun_out/c1_trace.txt 2>&1
timeout 600 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py quick > gpurun_out/c1_san_$tool.txt 2>&1; echo "rc=$?" >> gpurun_out/c1_san_$tool.txt
done
tail -3 gpurun_out/c1_pytest.txt; cat gpurun_out/c1_umma2.txt; tail -2 gpurun_out/c1_san_*.txt
