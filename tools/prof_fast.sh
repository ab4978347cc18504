# one full-set + source capture of k_fast (the big 128-image launch) -> gpurun_out/$1.ncu-rep
set -e
python bench.py --frames 256 --steps 1 --warmup 1 --no-cpu-baseline --parity-frames 0 > /dev/null 2> gpurun_out/$1_err.txt
ncu --set full --clock-control none --import-source on -k regex:k_fast -s 4 -c 1 -o gpurun_out/$1 -f \
  python bench.py --frames 256 --steps 1 --warmup 1 --no-cpu-baseline --parity-frames 0 > gpurun_out/$1_ncu.log 2>&1
ls -la gpurun_out/$1.ncu-rep
