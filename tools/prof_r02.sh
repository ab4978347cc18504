# Round-2 profile recipe (run on the GPU box): bench lines first (exit 0 without ncu), then the ncu launch list and one
# full-set capture of a chunk's kernels.  Outputs land in gpurun_out/ and are summarised into profiles/ here.
set -e
python bench.py > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_err.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_bench_reference.json 2>> gpurun_out/r02z_err.txt
python bench.py --workload 720p --frames 800 --no-cpu-baseline --parity-frames 64 > gpurun_out/r02z_bench_720p.json 2>> gpurun_out/r02z_err.txt
for b in 16 32; do python bench.py --workload 720p --frames 800 --batch $b --no-cpu-baseline --parity-frames 0 > gpurun_out/r02z_bench_720p_b$b.json 2>> gpurun_out/r02z_err.txt; done
python bench.py --mode classic --no-cpu-baseline > gpurun_out/r02z_bench_classic.json 2>> gpurun_out/r02z_err.txt
python bench.py --total-frames 20000 --no-cpu-baseline --parity-frames 64 > gpurun_out/r02z_strong_n1.json 2>> gpurun_out/r02z_err.txt
python bench.py --workload tracking --steps 3 --warmup 1 > gpurun_out/r02z_bench_tracking.json 2>> gpurun_out/r02z_err.txt
python bench.py --workload tracking --impl reference --steps 1 --warmup 1 > gpurun_out/r02z_bench_tracking_reference.json 2>> gpurun_out/r02z_err.txt
python bench.py --frames 256 --steps 1 --warmup 1 --no-cpu-baseline --parity-frames 0 > /dev/null 2>> gpurun_out/r02z_err.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/r02z_launches_ncu.csv \
  python bench.py --frames 256 --steps 1 --warmup 1 --no-cpu-baseline --parity-frames 0 > gpurun_out/r02z_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 24 -c 12 -o gpurun_out/r02z_full -f \
  python bench.py --frames 256 --steps 1 --warmup 1 --no-cpu-baseline --parity-frames 0 > gpurun_out/r02z_ncu2.log 2>&1
ls -la gpurun_out/r02z_full.ncu-rep
