#!/bin/bash
# GPU-box helper of round 2: full GPU test suite, then a short bench line; prints the headline figures
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py "$@" > gpurun_out/r02_bench_q.json 2> gpurun_out/r02_bench_q.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_q.json").read().strip().splitlines()[-1])
e = d["e2e"]
print("value", d["value"], "chunk1024", d.get("value_chunk_1024"), "e2e", e["value"], "e2e1024", e.get("value_chunk_1024"),
      e.get("chunk_1024_over_headline"), "launches", d["gpu_launches"])
if d.get("train_step"):
    t = d["train_step"]
    print({k: (v.get("ms_per_step") if isinstance(v, dict) else v) for k, v in t.items() if k in ("ms_per_step", "arena_adam_eager", "graphed", "graphed_arena_adam")})
PY
