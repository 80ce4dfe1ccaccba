#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "forward_lanes" --timeout 150 > gpurun_out/c30_lanes.txt 2>&1; echo "lanes rc=$?" >> gpurun_out/c30_lanes.txt; tail -n 8 gpurun_out/c30_lanes.txt
timeout 500 python bench.py --no-cpu-baseline --no-lib-baseline --no-eval > gpurun_out/c30_bench3.json 2> gpurun_out/c30_bench.err; tail -c 400 gpurun_out/c30_bench.err
timeout 500 python bench.py --lanes 1 --no-cpu-baseline --no-lib-baseline --no-eval > gpurun_out/c30_bench1.json 2>> gpurun_out/c30_bench.err
timeout 500 python bench.py --lanes 2 --no-cpu-baseline --no-lib-baseline --no-eval > gpurun_out/c30_bench2.json 2>> gpurun_out/c30_bench.err
timeout 300 python tools/ab_matrix.py stock '' 'pdl=0' '' 'pdl=0' > gpurun_out/c30_ab.txt 2>&1; grep forward gpurun_out/c30_ab.txt
python -c "
import json
for f in ('c30_bench3','c30_bench1','c30_bench2'):
    try:
        d=json.load(open('gpurun_out/%s.json'%f)); print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'single', round(d['single_stream']['value']), 'e2e', round(d['e2e']['value']), d['config']['in_flight'], d['clocks'])
    except Exception as e: print(f,'ERR',e)
"
