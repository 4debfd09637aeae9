O=gpurun_out/r2aa
mkdir -p $O
( time timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err ) 2>&1 | grep real; tail -3 $O/bench_default.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2aa/bench_default.json'))
print({k:b[k] for k in ('value','ms_per_step','ms_per_microbatch_256','tflops','clocks','gpu_launches')})
print('e2e', b['e2e']['value'], 'roofline', b['roofline']['frac'], b['roofline']['frac_of_composite_floor'])
print('parity', b.get('parity'))
ex=b['extra']
for k in ('fp16_nhwc','qs8_nhwc','qs8_nchw4'):
    print(k, {kk:ex[k][kk] for kk in ex[k] if kk in ('ms_per_microbatch','ms','tflops','tops','frac_of_int8_peak')})
for k in ex:
    if k.startswith('depthwise') or k.startswith('pointwise'):
        print(k, ex[k]['ms'], ex[k]['frac_hbm'])
for r in ex['adjacent_ops']: print(r)
print(b['cpu_baseline'])
PY
timeout 300 python bench.py --dtype qs8 --format nchw4 --global-batch 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-validate --layer-report $O/layers_nchw4.json > $O/bench_nchw4.json 2> $O/bench_nchw4.err; tail -2 $O/bench_nchw4.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2aa/layers_nchw4.json')); rows=d['layers'] if isinstance(d,dict) else d
tot=0
for r in rows:
    tot+=r['ms']*r['count']; print('%-10s x%d %s %.4f floor %.4f' % (r['name'], r['count'], r['algo'][-8:], r['ms'], r['floor_ms']))
print(tot)
PY
