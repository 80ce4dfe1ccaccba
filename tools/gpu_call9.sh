#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 120 python tools/cluster_dbg.py 4 > gpurun_out/c9_dbg4.txt 2>&1; echo "rc=$?" >> gpurun_out/c9_dbg4.txt; grep -v "^cl \|^ref " gpurun_out/c9_dbg4.txt | tail -n 24
timeout 120 python tools/cluster_dbg.py 4 8 224 224 > gpurun_out/c9_dbg4c.txt 2>&1; echo "rc=$?" >> gpurun_out/c9_dbg4c.txt; grep -v "^cl \|^ref " gpurun_out/c9_dbg4c.txt | tail -n 24
WIDTHS=pruned timeout 120 python tools/cluster_dbg.py 4 5 224 224 > gpurun_out/c9_dbg4p.txt 2>&1; echo "rc=$?" >> gpurun_out/c9_dbg4p.txt; grep -v "^ref " gpurun_out/c9_dbg4p.txt | tail -n 44
