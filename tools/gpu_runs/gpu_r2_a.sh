#!/bin/bash
# round 2, GPU call A: validate HEAD + first A/Bs
O=gpurun_out/r2a; mkdir -p $O
P=molecular-dynamics-simulation---lennard-jones-monoatomic-fluid_b200
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi.txt
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.log
LJMD_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_zz_gpu_experimental.py tests/test_zz_gpu_analysis.py -q > $O/pytest_exp.log 2>&1; echo "rc=$?" >> $O/pytest_exp.log
timeout 300 python bench.py --workload cfg4_nlist --steps 640 --warmup 128 --no-cpu > $O/cfg4_nlist.json 2> $O/cfg4_nlist.err
timeout 300 python bench.py --workload cfg3_allpairs --steps 100 --warmup 10 --no-cpu > $O/cfg3_allpairs.json 2> $O/cfg3_allpairs.err
LJMD_LIB_PATH=$PWD/$P/lib/libljmd_mic1.so timeout 300 python bench.py --workload cfg3_allpairs --steps 100 --warmup 10 --no-cpu > $O/cfg3_allpairs_mic1.json 2> $O/cfg3_allpairs_mic1.err
timeout 300 python bench.py --workload cfg2 --steps 20000 --warmup 2000 --no-cpu > $O/cfg2.json 2> $O/cfg2.err
LJMD_LIB_PATH=$PWD/$P/lib/libljmd_mic1.so timeout 300 python bench.py --workload cfg2 --steps 20000 --warmup 2000 --no-cpu > $O/cfg2_mic1.json 2> $O/cfg2_mic1.err
timeout 300 python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu > $O/cfg2_20.json 2> $O/cfg2_20.err
LJMD_COMPAT_GRAPH=1 timeout 300 python bench.py --workload cfg2 --steps 2000 --warmup 200 --no-cpu > $O/cfg2_compatgraph.json 2> $O/cfg2_compatgraph.err
timeout 300 python tools/bench_analysis.py > $O/analysis.jsonl 2> $O/analysis.err
LJMD_CORR_TILED=1 timeout 300 python tools/bench_analysis.py > $O/analysis_tiled.jsonl 2> $O/analysis_tiled.err
LJMD_SYM_DEBUG=1 timeout 120 python tools/warp_times.py > $O/warp_times.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lj_pairs_sym -c 1 -o $O/ncu_cfg3_allpairs python bench.py --workload cfg3_allpairs --steps 3 --warmup 3 --no-cpu > $O/ncu_cfg3.log 2>&1
LJMD_LIB_PATH=$PWD/$P/lib/libljmd_mic1.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:lj_pairs_sym -c 1 -o $O/ncu_cfg3_allpairs_mic1 python bench.py --workload cfg3_allpairs --steps 3 --warmup 3 --no-cpu > $O/ncu_cfg3_mic1.log 2>&1
ls -la $O
