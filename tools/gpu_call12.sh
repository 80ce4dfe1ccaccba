#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/ab_matrix.py stock 'FD_TC_CLUSTER=1' 'FD_TC_CLUSTER=1,FD_TC_EPI_HIGH=1' 'FD_TC_CLUSTER=1' 'FD_TC_CLUSTER=1,FD_TC_EPI_HIGH=1' > gpurun_out/c12_ab.txt 2>&1
cat gpurun_out/c12_ab.txt
timeout 400 python tools/ab_matrix.py pruned 'FD_TC_CLUSTER=1' 'FD_TC_CLUSTER=1,FD_TC_EPI_HIGH=1' > gpurun_out/c12_ab_pruned.txt 2>&1
cat gpurun_out/c12_ab_pruned.txt
