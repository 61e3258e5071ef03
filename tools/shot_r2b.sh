#!/bin/bash
# Last GPU call of round 2 (one B200): the GPU suite with the new thread / cross-stream tests, then the A/B of the late
# background-fill fork against the default on the same box, then the default bench line.
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q -rf --durations=8 > $out/r2b_pytest_gpu.txt 2>&1; echo "pytest exit $?"; tail -4 $out/r2b_pytest_gpu.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $out/r2b_bench_c2only.json 2> $out/r2b_bench_c2only.err; echo "c2-only exit $?"
PS_FILL_FORK_LATE=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > $out/r2b_bench_c2only_late_fork.json 2> $out/r2b_bench_c2only_late_fork.err; echo "late fork exit $?"
python bench.py --steps 10 --warmup 3 > $out/r2b_bench_default.json 2> $out/r2b_bench_default.err; echo "default exit $?"
