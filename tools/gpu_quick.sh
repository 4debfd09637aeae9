# quick GPU pass: parity tests + bf16 / qs8 bench with per-layer reports (no ncu)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/layers_bf16.json > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo bench=$?; tail -3 gpurun_out/bench_bf16.err
timeout 600 python bench.py --steps 20 --warmup 3 --dtype qs8 --no-cpu-baseline --layer-report gpurun_out/layers_qs8.json > gpurun_out/bench_qs8.json 2> gpurun_out/bench_qs8.err; echo benchq=$?; tail -3 gpurun_out/bench_qs8.err
python - <<'PY'
import json
for f in ['bf16','qs8']:
    try:
        b=json.load(open('gpurun_out/bench_%s.json'%f)); print(f, 'ms_per_step', b['ms_per_step'], 'TF', b['tflops'], 'e2e', b['e2e']['value'], b['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
