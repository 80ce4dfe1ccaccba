#!/bin/bash
# Regenerates the round-2 evidence under gpurun_out/ on a GPU box (then copied into profiles/ by hand, see profiles/README.md):
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/r02_evidence.sh'
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
B="python bench.py --lanes 1 --steps 2 --warmup 3 --graph 0 --no-cpu-baseline --no-lib-baseline --no-eval --stage-iters 1 --e2e-steps 6"   # one lane: the 15 launches of a forward in order
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_tc|block_tc|stem_tc" --launch-skip 30 --launch-count 30 --csv --log-file gpurun_out/g_launches.csv $B > gpurun_out/g_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chain_tc|block_tc|stem_tc" --launch-skip 30 --launch-count 15 -o gpurun_out/prof_r2g $B > gpurun_out/g_ncu_full.log 2>&1
timeout 200 python tools/trace_chain.py > gpurun_out/g_trace_chain.txt 2>&1
timeout 300 python tools/trace_stage.py 1 2 3 5 12 13 14 17 18 > gpurun_out/g_trace.txt 2>&1
for cfg in "--widths pruned" "--batch 32" "--dtype bf16" "--hw 480 640 --batch 16"; do
  timeout 300 python bench.py $cfg --no-cpu-baseline --no-lib-baseline --no-eval --e2e-steps 60 > "gpurun_out/g_bench_$(echo $cfg | tr -d ' -').json" 2>> gpurun_out/g_bench.err
done
for tool in memcheck synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py quick > gpurun_out/g_san_$tool.txt 2>&1; echo "rc=$?" >> gpurun_out/g_san_$tool.txt
done
# afterwards, here: python profiles/extract_ncu.py gpurun_out/prof_r2g.ncu-rep profiles/r02_final ; python profiles/stall_table.py gpurun_out/prof_r2g.ncu-rep profiles/r02_final_stalls.md r2
