#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
V=$PWD/fastdepth_b200/libfastdepth_b200_smallsmem.so
timeout 600 python -m pytest tests -m gpu -q -x --timeout 200 > gpurun_out/c27_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c27_pytest.txt; tail -n 4 gpurun_out/c27_pytest.txt
for r in 1 2; do
echo "== small budget (round-1 plans)" >> gpurun_out/c27_ab.txt
FD_B200_LIB=$V timeout 200 python tools/ab_matrix.py stock '' >> gpurun_out/c27_ab.txt 2>&1
echo "== main (227 KB budget)" >> gpurun_out/c27_ab.txt
timeout 200 python tools/ab_matrix.py stock '' >> gpurun_out/c27_ab.txt 2>&1
done
echo "== pruned small / main" >> gpurun_out/c27_ab.txt
FD_B200_LIB=$V timeout 200 python tools/ab_matrix.py pruned '' >> gpurun_out/c27_ab.txt 2>&1
timeout 200 python tools/ab_matrix.py pruned '' >> gpurun_out/c27_ab.txt 2>&1
cat gpurun_out/c27_ab.txt
