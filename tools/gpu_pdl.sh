set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl.json 2>> gpurun_out/bench_pdl.err
B200CB_NO_PDL=1 timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopdl.json 2>> gpurun_out/bench_pdl.err
python - <<'PY'
import json
for f in ['pdl','nopdl']:
    b=json.load(open('gpurun_out/bench_%s.json'%f)); print(f, 'ms_per_step %.4f TF %.1f e2e %.0f' % (b['ms_per_step'], b['tflops'], b['e2e']['value']), b['clocks'])
PY
done
timeout 300 python bench.py --steps 30 --warmup 3 --dtype qs8 --no-cpu-baseline | python -c "import json,sys; b=json.loads(sys.stdin.read()); print('qs8 ms_per_step', b['ms_per_step'])"
