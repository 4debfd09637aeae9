set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -s 162 -c 3 -f -o gpurun_out/prof_qs8 python bench.py --steps 2 --warmup 3 --dtype qs8 --no-cpu-baseline > gpurun_out/ncu_qs8.log 2>&1; echo ncu=$?
