#!/bin/bash
# A/B of libvslam_b200.so build variants on one box.  AB_LIST="name[@ENV=VAL[,ENV=VAL]] ..." ; libs in build/ab/lib_<name>.so
mkdir -p gpurun_out
for item in ${AB_LIST}; do
  n=${item%%@*}; e=""; [ "$item" != "$n" ] && e=$(echo "${item#*@}" | tr ',' ' ')
  tag=$(echo "$item" | tr '@=,' '___')
  env $e VSLAM_B200_LIB=$PWD/build/ab/lib_$n.so timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --parity-frames 64 > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python - "$tag" <<'PY'
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
