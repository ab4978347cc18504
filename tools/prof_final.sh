# Headline bench line + ncu launch list + one full-set capture of a chunk's kernels (the final code of the round)
set -e
python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_err.txt
python bench.py --frames 256 --steps 1 --warmup 1 --no-cpu-baseline --parity-frames 0 > /dev/null 2>> gpurun_out/r02f_err.txt
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv --log-file gpurun_out/r02f_launches_ncu.csv \
  python bench.py --frames 256 --steps 1 --warmup 1 --no-cpu-baseline --parity-frames 0 > gpurun_out/r02f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -s 24 -c 12 -o gpurun_out/r02f_full -f \
  python bench.py --frames 256 --steps 1 --warmup 1 --no-cpu-baseline --parity-frames 0 > gpurun_out/r02f_ncu2.log 2>&1
ls -la gpurun_out/r02f_full.ncu-rep
