#!/bin/bash
# One-GPU measurement set of a round (run under gpurun from the repository root); everything lands in gpurun_out/.
#   bash tools/final_runs_1gpu.sh <tag>
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 $out/${tag}_pytest.log
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err; echo "reference arm exit $?"
for wl in c1 c3 c5_3d c5_2d c4; do
  python bench.py --steps 5 --warmup 3 --workload $wl --no-cpu-baseline > $out/${tag}_bench_${wl}.json 2> $out/${tag}_bench_${wl}.err; echo "$wl exit $?"
done
python tools/latency_probe.py 3d > $out/${tag}_latency_3d.log 2>&1; head -2 $out/${tag}_latency_3d.log
python tools/latency_probe.py 2d > $out/${tag}_latency_2d.log 2>&1; head -2 $out/${tag}_latency_2d.log
python tools/bin_probe.py c2 1 4 fwdbwd 1 > $out/${tag}_single_view_stages.log 2>&1; tail -1 $out/${tag}_single_view_stages.log
# ncu: launch list of a short default bench run (per-launch times are cold-cache and serialised: compare shares)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $out/${tag}_ncu_plain.json 2> $out/${tag}_ncu_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_c2.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > $out/${tag}_ncu_launches.log 2>&1
echo "launch list exit $?"
