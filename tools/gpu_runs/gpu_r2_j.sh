#!/bin/bash
# 2-GPU validation of the slab-local rebuild
N=${1:-2}
O=gpurun_out/r2j_n$N; mkdir -p $O
P=$PWD/molecular-dynamics-simulation---lennard-jones-monoatomic-fluid_b200
timeout 600 python -m pytest tests/test_gpu_nlist.py -x -q > $O/pytest_nlist.log 2>&1; echo "rc=$?" >> $O/pytest_nlist.log
LJMD_SLAB_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_multi_process.py -x -q -k "$N]" > $O/pytest_mp_local.log 2>&1; echo "rc=$?" >> $O/pytest_mp_local.log
LJMD_SLAB_LOCAL=0 timeout 600 python -m pytest tests/test_gpu_multi_process.py -x -q -k "$N]" > $O/pytest_mp_global.log 2>&1; echo "rc=$?" >> $O/pytest_mp_global.log
run() { name=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 200 --warmup 20 > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name rc=$?" >> $O/runs.log; }
run local LJMD_X=0
run global LJMD_SLAB_LOCAL=0
run pub_local LJMD_LIB_PATH=$P/lib/libljmd_pub.so
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu --no-secondary > $O/cfg4_single.json 2> $O/cfg4_single.err
ls $O
