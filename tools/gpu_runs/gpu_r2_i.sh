#!/bin/bash
# multi-GPU validation of the graph'ed sharded chunk + timeline (run with gpurun --gpus N)
N=${1:-2}
O=gpurun_out/r2i_n$N; mkdir -p $O
P=$PWD/molecular-dynamics-simulation---lennard-jones-monoatomic-fluid_b200
LJMD_SLAB_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_multi_process.py -x -q -k "$N]" > $O/pytest_mp.log 2>&1; echo "rc=$?" >> $O/pytest_mp.log
LJMD_LIB_PATH=$P/lib/libljmd_pub.so timeout 600 python -m pytest tests/test_gpu_multi_process.py -x -q -k "$N]" > $O/pytest_mp_pub.log 2>&1; echo "rc=$?" >> $O/pytest_mp_pub.log
tl() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/slab_timeline.py 96 > $O/timeline_$name.txt 2>&1; }
tl base LJMD_X=0
tl pub LJMD_LIB_PATH=$P/lib/libljmd_pub.so
tl pub_nograph LJMD_LIB_PATH=$P/lib/libljmd_pub.so LJMD_SLAB_GRAPH=0
run() { name=$1; shift; env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 200 --warmup 20 > $O/bench_$name.json 2> $O/bench_$name.err; echo "$name rc=$?" >> $O/runs.log; }
run base LJMD_X=0
run pub LJMD_LIB_PATH=$P/lib/libljmd_pub.so
run pub_nograph LJMD_LIB_PATH=$P/lib/libljmd_pub.so LJMD_SLAB_GRAPH=0
run pub_driverlike LJMD_LIB_PATH=$P/lib/libljmd_pub.so
ls $O
