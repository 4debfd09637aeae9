# data collection: depthwise / pointwise tables, DRAM traffic of one bf16 step, full ncu captures
set -x
mkdir -p gpurun_out
for dt in bf16 qs8 f16 f32; do timeout 300 python tools/bench_layers.py --set dw --dtype $dt --out gpurun_out/dw_$dt.json > gpurun_out/dw_$dt.log 2>&1; tail -2 gpurun_out/dw_$dt.log; done
for dt in bf16 qs8; do timeout 300 python tools/bench_layers.py --set pw --dtype $dt --out gpurun_out/pw_$dt.json > gpurun_out/pw_$dt.log 2>&1; tail -2 gpurun_out/pw_$dt.log; done
# per-launch time + DRAM bytes of the 53 launches of one timed step (launches 0..158 are the 3 warm-up steps)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'igemm|fold|dwconv|transpose' -s 159 -c 53 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo ncu1=$?
# full captures: l3.bN.c1/c2/c3 region (pair tiles) = launches 33.. of the step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -s 192 -c 4 -f -o gpurun_out/prof_l3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l3.log 2>&1; echo ncu2=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv -s 40 -c 3 -f -o gpurun_out/prof_dw python tools/bench_layers.py --set dw --dtype bf16 --reps 4 > gpurun_out/ncu_dw.log 2>&1; echo ncu3=$?
ls -la gpurun_out | head -50
