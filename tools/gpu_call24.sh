#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/c24.txt
run() { echo "== $*" >> gpurun_out/c24.txt; env "$@" timeout 70 python tools/cfg_run.py stock $EXTRA >> gpurun_out/c24.txt 2>&1; echo "rc=$?" >> gpurun_out/c24.txt; }
EXTRA=""; run FD_TC_CLUSTER_MULTIWAVE=1 FD_TC_DW_TEAMS=1 FD_TC_CLUSTER=4
EXTRA=""; run FD_TC_CLUSTER_MULTIWAVE=1 FD_TC_DW_TEAMS=1 FD_TC_CLUSTER=2
EXTRA=""; run FD_TC_CLUSTER_MULTIWAVE=1 FD_TC_CLUSTER=4
grep -v "^Traceback\|^  File\|^    " gpurun_out/c24.txt | cut -c1-250 | sort | uniq -c | sort -rn | head -20
