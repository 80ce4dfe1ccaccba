#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tile_sharing" --timeout 200 > gpurun_out/c19_cl.txt 2>&1; echo "cl rc=$?" >> gpurun_out/c19_cl.txt; tail -n 6 gpurun_out/c19_cl.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "agree_bitwise and not tile_sharing" --timeout 300 > gpurun_out/c19_bitwise.txt 2>&1; echo "bitwise rc=$?" >> gpurun_out/c19_bitwise.txt; tail -n 6 gpurun_out/c19_bitwise.txt
if grep -q "rc=0" gpurun_out/c19_cl.txt && grep -q "rc=0" gpurun_out/c19_bitwise.txt; then
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 --deselect tests/test_gpu_parity.py::test_tile_sharing_clusters_agree_bitwise --deselect tests/test_gpu_parity.py::test_epilogue_organisations_and_item_shapes_agree_bitwise > gpurun_out/c19_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/c19_pytest.txt; tail -n 5 gpurun_out/c19_pytest.txt
fi
timeout 300 python tools/ab_matrix.py stock '' > gpurun_out/c19_ab.txt 2>&1; cat gpurun_out/c19_ab.txt
timeout 500 python bench.py > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err; tail -c 300 gpurun_out/c19_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c19_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['bound'], d['roofline']['frac'], d['roofline']['kernel'][:40])"
