set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'igemm|fold|dwconv' -s 162 -c 54 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo ncu1=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -s 159 -c ${NCU_COUNT:-6} -f -o gpurun_out/prof_igemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo ncu2=$?
