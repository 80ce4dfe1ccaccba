#!/bin/bash
# final evidence of HEAD: bench line, launch list, ncu full capture of one forward, timelines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; tail -c 300 gpurun_out/g_bench.err
B="python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline --no-lib-baseline --no-eval --stage-iters 1 --e2e-steps 6"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chain_tc|block_tc|stem_tc" --launch-skip 30 --launch-count 30 --csv --log-file gpurun_out/g_launches.csv $B > gpurun_out/g_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"chain_tc|block_tc|stem_tc" --launch-skip 30 --launch-count 15 -o gpurun_out/prof_r2g $B > gpurun_out/g_ncu_full.log 2>&1
ls -la gpurun_out/prof_r2g.ncu-rep
timeout 200 python tools/trace_chain.py > gpurun_out/g_trace_chain.txt 2>&1
timeout 300 python tools/trace_stage.py 1 2 3 5 12 13 14 17 18 > gpurun_out/g_trace.txt 2>&1
timeout 300 python bench.py --widths pruned --no-cpu-baseline --no-lib-baseline --no-eval --e2e-steps 60 > gpurun_out/g_bench_pruned.json 2>> gpurun_out/g_bench.err
python -c "
import json,glob
for f in sorted(glob.glob('gpurun_out/g_bench*.json')):
    try:
        d=json.load(open(f)); print(f, round(d['value']), round(d['e2e']['value']), d['roofline']['bound'], round(d['roofline']['frac'],3), d['roofline']['traffic'], d['parity']['max_rel_err_vs_oracle'])
    except Exception as e: print(f, 'ERR', e)
"
