python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for v in "VSLAM_BRIEF_TMA=0" "VSLAM_BRIEF_TMA=1" "VSLAM_B200_LIB=$PWD/orb-slam3-rust_b200/vslam_b200/libvslam_tma80.so"; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --parity-frames 16 > gpurun_out/r02j.json 2>gpurun_out/r02j_err.txt || tail -3 gpurun_out/r02j_err.txt
  python - <<PY
import json
d=json.load(open("gpurun_out/r02j.json"))
print("$v", d["value"], d["e2e"]["value"], d["p50_ms_single_pair_latency"], d["roofline"]["stage_ms_per_step"]["k_brief"], d["parity"]["mismatching_frames"])
PY
done
