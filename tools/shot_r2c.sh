#!/bin/bash
# Round 2, last GPU call: the instruction diet of block_lists / raster_fwd6 / transpose32 -- GPU suite first, the bench only
# if it is green.
out=gpurun_out
mkdir -p $out
timeout 150 python -m pytest tests -m gpu -q -rf > $out/r2c_pytest_gpu.txt 2>&1; rc=$?; echo "pytest exit $rc"; tail -4 $out/r2c_pytest_gpu.txt
if [ $rc -ne 0 ]; then exit 0; fi
python bench.py --steps 10 --warmup 3 > $out/r2c_bench_default.json 2> $out/r2c_bench_default.err; echo "default exit $?"
python -c "
import json
d = json.loads(open('$out/r2c_bench_default.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), {k: round(v, 3) for k, v in d['stage_ms_per_step'].items()})
print('c3', round(d['also']['c3']['value']), {k: round(v, 3) for k, v in d['also']['c3']['stage_ms_per_step'].items()})
"
