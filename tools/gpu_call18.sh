#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/c18.txt
run() { echo "== $*" >> gpurun_out/c18.txt; env "$@" timeout 60 python tools/cfg_run.py stock $EXTRA >> gpurun_out/c18.txt 2>&1; echo "rc=$?" >> gpurun_out/c18.txt; }
EXTRA=""; run FD_TC_CLUSTER=4
run FD_TC_CLUSTER=2
grep -v "^Traceback\|^  File\|^    " gpurun_out/c18.txt | cut -c1-200
if grep -q "rc=124" gpurun_out/c18.txt; then echo "STILL HANGS"; else
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "agree_bitwise" --timeout 300 > gpurun_out/c18_bitwise.txt 2>&1; echo "bitwise rc=$?" >> gpurun_out/c18_bitwise.txt; tail -n 5 gpurun_out/c18_bitwise.txt
fi
timeout 500 python tools/ab_matrix.py stock 'FD_TC_DW_TEAMS=1' '' 'FD_TC_DW_TEAMS=1' '' > gpurun_out/c14_ab.txt 2>&1
cat gpurun_out/c14_ab.txt
timeout 300 python tools/ab_matrix.py pruned 'FD_TC_DW_TEAMS=1' '' > gpurun_out/c14_ab_pruned.txt 2>&1
cat gpurun_out/c14_ab_pruned.txt
