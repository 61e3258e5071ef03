#!/bin/bash
# ncu --set full of every kernel of one c2 / c3 step at the bench batch sizes (one ncu use per gpurun call).
#   bash tools/final_runs_ncu_full.sh <tag> <workload> <frames> <matching launches per repetition: 10 for 3D, 8 for 2D>
tag=${1:-r2}; wl=${2:-c2}; frames=${3:-256}; per=${4:-10}
out=gpurun_out
mkdir -p $out
python tools/bin_probe.py $wl $frames 3 > $out/${tag}_probe_${wl}.log 2>&1 && cat $out/${tag}_probe_${wl}.log &&
ncu --set full --clock-control none --import-source on \
    -k regex:"project|depth_rank|partition|sort_lists|block_lists|raster_fwd|raster_bwd|fill_empty" -s $((2 * per)) -c $per \
    -o $out/${tag}_full_${wl} python tools/bin_probe.py $wl $frames 3 > $out/${tag}_ncu_full_${wl}.log 2>&1
echo "ncu full exit $?"; tail -2 $out/${tag}_ncu_full_${wl}.log
