set -x
mkdir -p gpurun_out
timeout 120 python tools/membw.py > gpurun_out/membw.json 2>&1; cat gpurun_out/membw.json
timeout 300 python -m pytest tests -m gpu -x -q -k "selftest or stem or preprocess" 2>&1 | tail -3
for dt in bf16 qs8 f16 f32; do timeout 300 python tools/bench_layers.py --set dw --dtype $dt --out gpurun_out/dw_$dt.json > gpurun_out/dw_$dt.log 2>&1; tail -1 gpurun_out/dw_$dt.log; done
for dt in bf16 qs8; do timeout 300 python tools/bench_layers.py --set pw --dtype $dt --out gpurun_out/pw_$dt.json > gpurun_out/pw_$dt.log 2>&1; tail -1 gpurun_out/pw_$dt.log; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dwconv -s 40 -c 4 -f -o gpurun_out/prof_dw python tools/bench_layers.py --set dw --dtype bf16 --reps 4 > gpurun_out/ncu_dw.log 2>&1; echo ncu=$?
