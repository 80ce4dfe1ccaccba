#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for m in w2 w4; do
timeout 120 python tools/cluster_dbg.py $m 8 224 224 > gpurun_out/c11_dbg_$m.txt 2>&1; echo "rc=$?" >> gpurun_out/c11_dbg_$m.txt; grep "^cl \|OK\|MISM\|DIFF\|rror" gpurun_out/c11_dbg_$m.txt | tail -n 30
done
timeout 600 python tools/ab_matrix.py stock 'FD_TC_CLUSTER=1,FD_TC_WMC=1' 'FD_TC_CLUSTER=1,FD_TC_WMC=2' 'FD_TC_CLUSTER=1,FD_TC_WMC=4' 'chain=0,FD_TC_CLUSTER=1,FD_TC_WMC=4' > gpurun_out/c11_ab.txt 2>&1
cat gpurun_out/c11_ab.txt
timeout 400 python tools/ab_matrix.py pruned 'FD_TC_CLUSTER=1,FD_TC_WMC=1' 'FD_TC_CLUSTER=1,FD_TC_WMC=2' 'FD_TC_CLUSTER=1,FD_TC_WMC=4' > gpurun_out/c11_ab_pruned.txt 2>&1
cat gpurun_out/c11_ab_pruned.txt
