#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/cluster_dbg.py 4 8 224 224 > gpurun_out/c10_dbg4.txt 2>&1; echo "rc=$?" >> gpurun_out/c10_dbg4.txt; tail -n 2 gpurun_out/c10_dbg4.txt
timeout 600 python tools/ab_matrix.py stock 'FD_TC_CLUSTER=1' 'FD_TC_CLUSTER=4' 'FD_TC_CLUSTER=2' 'chain=0,FD_TC_CLUSTER=4' > gpurun_out/c10_ab.txt 2>&1
cat gpurun_out/c10_ab.txt
CHAIN=0 FD_TC_CLUSTER=4 timeout 300 python tools/trace_stage.py 7 13 14 > gpurun_out/c10_trace_cl4.txt 2>&1; cut -c1-230 gpurun_out/c10_trace_cl4.txt
