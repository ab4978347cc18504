#!/bin/bash
# Sweep of environment knobs / bench arguments on one box.  Each line of $1: "name | ENV=VAL ... | extra bench args"
mkdir -p gpurun_out
while IFS='|' read -r name envs args; do
  name=$(echo $name); [ -z "$name" ] && continue
  env $envs timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --parity-frames 0 $args > gpurun_out/abe_$name.json 2> gpurun_out/abe_$name.err
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/abe_{n}.json"))
    print(n, d["value"], d["e2e"]["value"], d["roofline"].get("launch_ms"), flush=True)
except Exception as e:
    print(n, "failed", e, flush=True)
PY
done < "$1"
