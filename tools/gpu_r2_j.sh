O=gpurun_out/r2j
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 5 --no-cpu-baseline --no-extras --layer-report $O/layers_bf16.json > $O/bench_bf16.json 2> $O/bench_bf16.err; echo bench=$?; tail -3 $O/bench_bf16.err
timeout 600 python bench.py --steps 5 --dtype qs8 --no-cpu-baseline --no-extras --layer-report $O/layers_qs8.json > $O/bench_qs8.json 2> $O/bench_qs8.err; echo bench=$?; tail -3 $O/bench_qs8.err
python - <<'PY'
import json
for k in ('bf16','qs8'):
    b=json.load(open('gpurun_out/r2j/bench_%s.json'%k))
    print(k, {x:b[x] for x in ('value','ms_per_microbatch_256','tflops')}, b['clocks']['sm_mhz'], b['parity']['ok'], 'e2e', b['e2e']['value'])
    L=json.load(open('gpurun_out/r2j/layers_%s.json'%k))['layers']
    print('  conv1', [r for r in L if r['name']=='conv1'][0]['ms'], 'sum', sum(r['ms']*r['count'] for r in L))
PY
