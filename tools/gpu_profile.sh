#!/bin/bash
# round-end evidence: default bench, ncu launch list and ncu --set full of one step (each ncu pass only after the plain run exited 0)
tag=${1:-s3}
set -x
python bench.py > gpurun_out/bench_default_$tag.json 2> gpurun_out/bench_default_$tag.err || { tail -c 2000 gpurun_out/bench_default_$tag.err; exit 1; }
python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu --no-mapfusion --no-bow --no-other-configs > gpurun_out/bench_b64_$tag.json 2>/dev/null || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_b64_$tag.csv \
    python bench.py --batch 64 --steps 2 --warmup 3 --no-cpu --no-mapfusion --no-bow --no-other-configs > /dev/null 2>&1
timeout 800 ncu --set full --clock-control none --import-source on --launch-skip 42 -c 14 -o gpurun_out/prof_full_$tag \
    python bench.py --batch 64 --steps 1 --warmup 3 --no-cpu --no-mapfusion --no-bow --no-other-configs > gpurun_out/ncu_full_$tag.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:fast_cells --launch-skip 3 -c 1 -o gpurun_out/prof_fast_b512_$tag \
    python bench.py --batch 512 --steps 1 --warmup 3 --no-cpu --no-mapfusion --no-bow --no-other-configs > /dev/null 2>&1
ls -la gpurun_out/ | tail -12
