#!/bin/bash
# round 2, GPU call F (1 GPU): sorted kernel v4 (8-chunk items, flag-driven finalize), cfg4 parity, sharded-variant overhead on one GPU
O=gpurun_out/r2f; mkdir -p $O
P=$PWD/molecular-dynamics-simulation---lennard-jones-monoatomic-fluid_b200
timeout 900 python -m pytest tests/test_gpu_sorted_allpairs.py tests/test_gpu_cfg4_parity.py tests/test_gpu_nlist.py -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python bench.py --workload cfg3_allpairs --steps 1000 --warmup 10 --no-cpu > $O/cfg3_allpairs.json 2> $O/cfg3_allpairs.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_cfg3.csv python bench.py --workload cfg3_allpairs --steps 8 --warmup 3 --no-cpu > $O/launches_cfg3.log 2>&1
# the sharded instantiation of the step kernel on ONE rank owning everything: its own overhead, without any peer
LJMD_BENCH_FORCE_SLABS=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu > $O/cfg4_slab1.json 2> $O/cfg4_slab1.err
LJMD_BENCH_FORCE_SLABS=1 LJMD_LIB_PATH=$P/lib/libljmd_pub.so timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu > $O/cfg4_slab1_pub.json 2> $O/cfg4_slab1_pub.err
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --no-secondary > $O/cfg4_single.json 2> $O/cfg4_single.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lj_pairs_sorted -c 1 -o $O/ncu_cfg3_sorted python bench.py --workload cfg3_allpairs --steps 3 --warmup 3 --no-cpu > $O/ncu_cfg3.log 2>&1
ls -la $O
