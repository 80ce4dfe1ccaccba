#!/bin/bash
# session-2 call #2: r02 evidence of the current code (launch list, ncu full capture of one forward, stage timelines, sanitizer)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline --no-lib-baseline --no-eval --stage-iters 1 --e2e-steps 6"
timeout 300 python tools/trace_stage.py 1 2 3 12 13 14 15 16 17 18 > gpurun_out/c6_trace.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_tc|block_tc|stem_tc" --launch-skip 30 --launch-count 30 --csv --log-file gpurun_out/c6_launches.csv $B > gpurun_out/c6_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chain_tc|block_tc|stem_tc" --launch-skip 30 --launch-count 15 -o gpurun_out/prof_r2a $B > gpurun_out/c6_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
for tool in memcheck synccheck racecheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py quick > gpurun_out/c6_san_$tool.txt 2>&1; echo "rc=$?" >> gpurun_out/c6_san_$tool.txt
  tail -n 4 gpurun_out/c6_san_$tool.txt
done
