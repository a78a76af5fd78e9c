#!/bin/bash
# parity subset + short bench with per-stage times; run under gpurun from the repo root
tag=${1:-check}
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-mapfusion --no-bow > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || tail -c 800 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print("value %.0f e2e %.0f ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
print({k: round(v["ms"],4) for k,v in d["roofline"]["stages"].items()})
PY
