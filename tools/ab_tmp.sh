python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for v in "VSLAM_SELECT_FUSED=0" "VSLAM_SELECT_FUSED=1" "VSLAM_SEL2_CAP=64" "VSLAM_SEL1_CAP=1024"; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --parity-frames 16 > gpurun_out/r02i_$v.json 2>gpurun_out/r02i_err.txt || tail -3 gpurun_out/r02i_err.txt
  python - <<PY
import json
d=json.load(open("gpurun_out/r02i_$v.json"))
print("$v", d["value"], d["e2e"]["value"], d["p50_ms_single_pair_latency"], d["roofline"]["stage_ms_per_step"]["k_select x8 (retainBest+Harris+angle)"], d["parity"]["mismatching_frames"])
PY
done
VSLAM_SELECT_FUSED=1 python tools/lat_probe.py
VSLAM_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_select -c 100 --csv --log-file gpurun_out/r02i_lat_launches.csv python tools/lat_probe.py > gpurun_out/r02i_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02i_lat_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
agg=collections.OrderedDict()
for r in rows[1:]:
    k=(r[ki][:40], r[gi]); agg.setdefault(k,[]).append(float(r[vi].replace(",","")))
for k,v in agg.items():
    v=sorted(v); print(k, len(v), "median us %.1f"%(v[len(v)//2]/1000))
PY
