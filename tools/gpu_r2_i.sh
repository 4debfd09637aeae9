# call I: new bench.py (self-validating, 8192-image sharded job, extras) + the new GPU tests
O=gpurun_out/r2i
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -5 $O/pytest_gpu.log
( time timeout 900 python bench.py --layer-report $O/layers_bf16.json > $O/bench_default.json 2> $O/bench_default.err ) 2> $O/bench_default.time; echo bench=$?; tail -3 $O/bench_default.err; cat $O/bench_default.time
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo ref=$?
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2i/bench_default.json'))
print({k:b[k] for k in ('value','ms_per_step','ms_per_microbatch_256','tflops','wall_ms_per_step','host','clocks')})
print('e2e', b.get('e2e'))
print('parity', b.get('parity'))
ex=b.get('extra',{})
for k,v in ex.items():
    if isinstance(v,dict): print(k, {kk:vv for kk,vv in v.items() if kk!='rows'})
    elif k!='layers_bf16': print(k, v)
print('cpu', b.get('cpu_baseline'))
PY
