#!/bin/bash
# A/B of k_fast build variants on one box: tools/ab_fast.sh (libs in build/ab/lib_<name>.so)
mkdir -p gpurun_out
for n in ${AB_LIST:-base T G W TG TGW base TGW}; do
  VSLAM_B200_LIB=$PWD/build/ab/lib_$n.so timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --parity-frames 64 > gpurun_out/ab_$n.json 2> gpurun_out/ab_$n.err
  python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/ab_{n}.json"))
    r=d["roofline"]
    print(n, d["value"], d["e2e"]["value"], d["p50_ms_single_pair_latency"], r.get("launch_ms"), r["achieved"], d["parity"]["mismatching_frames"], flush=True)
except Exception as e:
    print(n, "failed", e, flush=True)
PY
done
