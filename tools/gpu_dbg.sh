set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/layers_bf16.json > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; tail -2 gpurun_out/bench_bf16.err
timeout 300 python bench.py --steps 20 --warmup 3 --dtype qs8 --no-cpu-baseline --layer-report gpurun_out/layers_qs8.json > gpurun_out/bench_qs8.json 2> gpurun_out/bench_qs8.err; tail -2 gpurun_out/bench_qs8.err
B200CB_LIB=megengine_b200/lib/libb200cb_dbg.so timeout 300 python tools/dbg_timing.py bf16 > gpurun_out/dbg_bf16.log 2>&1; cat gpurun_out/dbg_bf16.log
python - <<'PY'
import json
for f in ['bf16','qs8']:
    b=json.load(open('gpurun_out/bench_%s.json'%f)); print(f, 'ms_per_step', b['ms_per_step'], 'TF', b['tflops'], 'e2e', b['e2e']['value'], b['clocks'])
PY
