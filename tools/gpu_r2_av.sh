O=gpurun_out/r2av
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -2 $O/pytest_gpu.log
for dt in bf16 f16; do timeout 300 python tools/bench_layers.py --set pw --dtype $dt --out $O/pw_$dt.json > $O/pw_$dt.log 2>&1; tail -1 $O/pw_$dt.log; done
python - <<'PY'
import json
r=json.load(open('gpurun_out/r2av/pw_bf16.json'))
for y in r['rows']: print('   %-40s %s %.4f (%.2f)'%(y['desc'].split(' d1')[0].split('n256 ')[1], y['algo'][-7:], y['ms'], y['frac_hbm']))
PY
timeout 600 python bench.py --global-batch 2048 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --layer-report $O/layers_bf16.json > $O/bench_bf16_short.json 2> $O/bench_bf16_short.err; echo bench=$?; tail -2 $O/bench_bf16_short.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2av/bench_bf16_short.json')); print('bf16 ms/mb %.4f value %.0f e2e %.0f' % (b['ms_per_microbatch_256'], b['value'], b['e2e']['value']), b['clocks'])
PY
