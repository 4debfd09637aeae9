set -x
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'igemm|fold|transpose' -s 477 -c 159 --csv --log-file gpurun_out/launches_nchw4.csv python bench.py --dtype qs8 --format nchw4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_nchw4.log 2>&1; echo ncu=$?
