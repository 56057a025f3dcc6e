#!/bin/bash
# Drop-in call at N = 2^20 with the position download overlapped with the step kernel: A/B parity + e2e.
O=gpurun_out/r2s; mkdir -p $O
timeout 200 python -m pytest tests/test_zz_gpu_ab_switches.py tests/test_gpu_nlist.py -k "overlap or page_locked or drop_in" -x -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 120 python bench.py --steps 20 --warmup 5 --no-secondary --cpu-budget 1 --long-steps 0 > $O/bench_overlap.json 2> $O/bench_overlap.err; echo "rc=$?" >> $O/bench_overlap.err
LJMD_COMPAT_OVERLAP=0 timeout 120 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu --long-steps 0 > $O/bench_serial.json 2> $O/bench_serial.err; echo "rc=$?" >> $O/bench_serial.err
tail -3 $O/pytest.log; python - <<'P'
import json
for f in ("bench_overlap", "bench_serial"):
    try:
        d = json.loads(open("gpurun_out/r2s/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["e2e"]["ms_per_step"], d["e2e"]["value"], d["ms_per_step"])
    except Exception as e:
        print(f, "failed", e)
P
