set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -4 gpurun_out/pytest_gpu.log
B200CB_PAIR=1 timeout 900 python -m pytest tests -m gpu -x -q -k "not dropin" > gpurun_out/pytest_pair.log 2>&1; echo pytest_pair=$?; tail -3 gpurun_out/pytest_pair.log
timeout 600 python bench.py --no-cpu-baseline --layer-report gpurun_out/layers_bf16.json > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo rc=$?
timeout 600 python bench.py --dtype qs8 --no-cpu-baseline --layer-report gpurun_out/layers_qs8.json > gpurun_out/bench_qs8.json 2> gpurun_out/bench_qs8.err; echo rc=$?
timeout 120 python tools/bench_layers.py --set pw --dtype bf16 --out gpurun_out/pw_bf16.json > gpurun_out/pw_bf16.log 2>&1; tail -1 gpurun_out/pw_bf16.log
timeout 120 python tools/bench_layers.py --set pw --dtype qs8 --out gpurun_out/pw_qs8.json > gpurun_out/pw_qs8.log 2>&1; tail -1 gpurun_out/pw_qs8.log
python - <<'PY'
import json
for f in ['bf16','qs8']:
    b=json.load(open('gpurun_out/bench_%s.json'%f)); print(f, 'ms_per_step %.4f TF %.1f e2e %.0f' % (b['ms_per_step'], b['tflops'], b['e2e']['value']), b['clocks'])
PY
