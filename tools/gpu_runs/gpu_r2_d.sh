#!/bin/bash
# round 2, GPU call D: sorted kernel v3 (cost-ordered items) + cfg4 parity test
O=gpurun_out/r2d; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sorted_allpairs.py tests/test_gpu_cfg4_parity.py tests/test_abi.py -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python bench.py --workload cfg3_allpairs --steps 100 --warmup 10 --no-cpu > $O/cfg3_allpairs.json 2> $O/cfg3_allpairs.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lj_pairs_sorted -c 1 -o $O/ncu_cfg3_sorted python bench.py --workload cfg3_allpairs --steps 3 --warmup 3 --no-cpu > $O/ncu_cfg3.log 2>&1
ls -la $O
