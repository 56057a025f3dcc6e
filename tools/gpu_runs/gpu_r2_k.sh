#!/bin/bash
# round 2, second 8-GPU call: slab-local rebuild at 4 / 8 ranks (bitwise tests), per-step timeline, strong scaling, cfg5 ensemble
O=gpurun_out/r2k_n8; mkdir -p $O
P=$PWD/molecular-dynamics-simulation---lennard-jones-monoatomic-fluid_b200
LJMD_SLAB_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_multi_process.py -x -q -k "4] or 8]" > $O/pytest_mp.log 2>&1; echo "rc=$?" >> $O/pytest_mp.log
tl() { name=$1; n=$2; shift; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 tools/slab_timeline.py 96 > $O/timeline_${name}_n$n.txt 2>&1; }
tl base 8 LJMD_X=0
tl pub 8 LJMD_LIB_PATH=$P/lib/libljmd_pub.so
run() { name=$1; n=$2; shift; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $n --steps 200 --warmup 20 > $O/bench_${name}_n$n.json 2> $O/bench_${name}_n$n.err; echo "$name n=$n rc=$?" >> $O/runs.log; }
run base 8 LJMD_X=0
run pub 8 LJMD_LIB_PATH=$P/lib/libljmd_pub.so
run pub_global 8 LJMD_LIB_PATH=$P/lib/libljmd_pub.so LJMD_SLAB_LOCAL=0
run pub 4 LJMD_LIB_PATH=$P/lib/libljmd_pub.so
env LJMD_LIB_PATH=$P/lib/libljmd_pub.so timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_pub_driverlike_n8.json 2> $O/bench_pub_driverlike_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --workload cfg5_ensemble --steps 2000 --warmup 200 > $O/bench_cfg5_n8.json 2> $O/bench_cfg5_n8.err
ls $O
