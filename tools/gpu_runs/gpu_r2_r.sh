#!/bin/bash
# Dress rehearsal of the round-end driver sequence at HEAD: the full GPU suite (incl. the tests against the
# reference-output fixture of the analysis kernels), smoke(), the default bench line.
O=gpurun_out/${OUT:-r2r}; mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q --durations=12 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 240 python bench.py --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; echo "rc=$?" >> $O/bench_default.err
tail -4 $O/pytest.log; tail -2 $O/smoke.log; cut -c1-400 $O/bench_default.json
