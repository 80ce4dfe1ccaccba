#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python tools/ab_matrix.py stock 'FD_TC_DW_TEAMS=1' '' 'FD_TC_DW_TEAMS=1' '' > gpurun_out/c14_ab.txt 2>&1
cat gpurun_out/c14_ab.txt
timeout 600 python tools/ab_matrix.py pruned 'FD_TC_DW_TEAMS=1' '' > gpurun_out/c14_ab_pruned.txt 2>&1
cat gpurun_out/c14_ab_pruned.txt
timeout 300 python tools/trace_stage.py 1 3 5 13 18 > gpurun_out/c14_trace.txt 2>&1; grep -E "^==|period|dw math|epilogue duration" gpurun_out/c14_trace.txt
