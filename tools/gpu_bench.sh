set -x
mkdir -p gpurun_out
timeout 600 python bench.py --layer-report gpurun_out/layers_bf16.json > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo rc=$?; cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 600 python bench.py --dtype qs8 --no-cpu-baseline --layer-report gpurun_out/layers_qs8.json > gpurun_out/bench_qs8.json 2> gpurun_out/bench_qs8.err; echo rc=$?; python -c "
import json; b=json.load(open('gpurun_out/bench_qs8.json')); print('qs8', b['ms_per_step'], b['e2e'])"
