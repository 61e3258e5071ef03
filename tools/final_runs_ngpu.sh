#!/bin/bash
# N-GPU measurement set (run under `gpurun --gpus N` from the repository root):  bash tools/final_runs_ngpu.sh <tag> <N>
tag=${1:-r2}; n=${2:-2}
out=gpurun_out
mkdir -p $out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29520"
$T bench.py --gpus $n --steps 10 --warmup 3 > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err; echo "default exit $?"
$T bench.py --gpus $n --steps 8 --warmup 3 --workload c4 > $out/${tag}_bench_${n}gpu_c4.json 2> $out/${tag}_bench_${n}gpu_c4.err; echo "c4 exit $?"
$T bench.py --gpus $n --steps 5 --warmup 3 --workload c5_3d --split-frames --fused-reduce --frames 64 --no-cpu-baseline > $out/${tag}_bench_${n}gpu_c5_3d_split_fused.json 2> $out/${tag}_c5f.err; echo "c5 fused exit $?"
$T bench.py --gpus $n --steps 5 --warmup 3 --workload c5_3d --split-frames --frames 64 --no-cpu-baseline > $out/${tag}_bench_${n}gpu_c5_3d_split_nccl.json 2> $out/${tag}_c5n.err; echo "c5 nccl exit $?"
$T bench.py --gpus $n --steps 5 --warmup 3 --workload c5_2d --no-cpu-baseline > $out/${tag}_bench_${n}gpu_c5_2d.json 2> $out/${tag}_c52d.err; echo "c5 2d exit $?"
if [ "$n" = "2" ]; then python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -2; fi
