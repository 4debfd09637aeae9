O=gpurun_out/r2ag
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "int8 or qs8 or nchw4 or bit_exact or resnet50 or selftest or dw or depthwise or full_size or graph or dropin" > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 $O/pytest_gpu.log
timeout 300 python bench.py --dtype qs8 --global-batch 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --layer-report $O/layers_qs8.json > $O/bench_qs8.json 2> $O/bench_qs8.err; echo bench=$?; tail -2 $O/bench_qs8.err
python -c "
import json; b=json.load(open('$O/bench_qs8.json')); print('qs8 ms/microbatch', b.get('ms_per_microbatch_256'), b['clocks'])
d=json.load(open('$O/layers_qs8.json')); rows=d if isinstance(d,list) else d['layers']
for r in rows:
    print('%-10s x%d %s %.4f floor %.4f' % (r['name'], r['count'], r['algo'][-8:], r['ms'], r['floor_ms']))
"
timeout 120 python tools/bench_layers.py --set dw --dtype qs8 --out $O/dw_qs8.json > $O/dw_qs8.log 2>&1; tail -12 $O/dw_qs8.log | cut -c1-200
timeout 120 python tools/bench_layers.py --set pw --dtype qs8 --out $O/pw_qs8.json > $O/pw_qs8.log 2>&1; tail -10 $O/pw_qs8.log | cut -c1-200
timeout 300 python bench.py --dtype qs8 --format nchw4 --global-batch 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-validate > $O/bench_nchw4.json 2> $O/bench_nchw4.err; python -c "
import json; b=json.load(open('$O/bench_nchw4.json')); print('nchw4 ms/microbatch', b.get('ms_per_microbatch_256'))"
