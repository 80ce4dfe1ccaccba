#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/c15_cfg.txt
run() { echo "== $*" >> gpurun_out/c15_cfg.txt; env "$@" timeout 90 python tools/cfg_run.py stock $EXTRA >> gpurun_out/c15_cfg.txt 2>&1; echo "rc=$?" >> gpurun_out/c15_cfg.txt; }
EXTRA=""
run FD_TC_DW_TEAMS=1 FD_TC_CLUSTER=1
run FD_TC_DW_TEAMS=1
run FD_TC_DW_TEAMS=1 FD_TC_MAX_NCTA=256 FD_TC_NO_COLSPLIT=1 FD_TC_NO_WIDE=1 FD_TC_CLUSTER=1
EXTRA="wait_sleep_ns=200"
run FD_TC_DW_TEAMS=1 FD_TC_MAX_NCTA=128
EXTRA=""
run FD_TC_DW_TEAMS=1 FD_TC_CLUSTER=2
run FD_TC_DW_TEAMS=1 FD_TC_CLUSTER=4
run FD_TC_DW_TEAMS=1 FD_TC_CLUSTER=1 FD_TC_WMC=2
run FD_TC_DW_TEAMS=1 FD_TC_CLUSTER=1 FD_TC_WMC=4
run FD_TC_CLUSTER=1
run FD_TC_CLUSTER=1 FD_TC_WMC=4
grep -v "^Traceback\|^  File\|^    " gpurun_out/c15_cfg.txt | cut -c1-400
