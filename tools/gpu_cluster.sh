# cluster-size experiment: parity under forced cluster sizes, then per-layer timings for C = 1, 2, 4
set -x
mkdir -p gpurun_out
for c in 2 4; do
  B200CB_CLUSTER=$c timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_c$c.log 2>&1; echo pytest_c$c=$?; tail -4 gpurun_out/pytest_c$c.log
done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
for dt in bf16 qs8; do
for c in 1 2 4; do
  B200CB_CLUSTER=$c timeout 300 python bench.py --steps 10 --warmup 3 --dtype $dt --no-cpu-baseline --layer-report gpurun_out/layers_${dt}_c$c.json > gpurun_out/bench_${dt}_c$c.json 2> gpurun_out/bench_${dt}_c$c.err; echo bench_$dt_c$c=$?; tail -2 gpurun_out/bench_${dt}_c$c.err
done
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layer-report gpurun_out/layers_bf16.json > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_*.json')):
    try:
        b=json.load(open(f)); print(f, 'ms_per_step %.3f TF %.1f' % (b['ms_per_step'], b['tflops']))
    except Exception as e: print(f, 'ERR', e)
PY
