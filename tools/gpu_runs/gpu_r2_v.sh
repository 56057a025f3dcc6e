#!/bin/bash
# Build kernel with the running append pointer: bit-exact list tests + long run (rebuilds amortised) + per-kernel time of a rebuild.
O=gpurun_out/r2v; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_nlist.py tests/test_gpu_cfg4_parity.py -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 120 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu --e2e-steps 3 > $O/bench.json 2> $O/bench.err; echo "rc=$?" >> $O/bench.err
LJMD_NL_NOGRAPH=1 STEPS=1 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nl_build_kernel -c 1 python tools/quick_nlist.py cfg4_nlist 2>&1 | grep -E "nl_build_kernel|gpu__time_duration" > $O/build_time.txt
tail -3 $O/pytest.log; cat $O/build_time.txt; python - <<'P'
import json
d = json.loads(open("gpurun_out/r2v/bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["long_run"])
P
