set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -s 159 -c 1 -f -o gpurun_out/prof_conv1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv1.log 2>&1; echo ncu=$?
