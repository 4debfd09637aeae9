O=gpurun_out/r2bl
mkdir -p $O
timeout 900 python bench.py --steps 20 --warmup 3 --layer-report $O/layers_bf16.json > $O/bench_bf16.json 2> $O/bench_bf16.err; echo bench=$?
timeout 600 python bench.py --dtype qs8 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --layer-report $O/layers_qs8.json > $O/bench_qs8.json 2> $O/bench_qs8.err; echo bench_qs8=$?
python - <<'PY'
import json
for f in ['bf16','qs8']:
    b=json.load(open('gpurun_out/r2bl/bench_%s.json'%f)); print(f, 'ms/mb %.4f TF %.1f value %.0f e2e %.0f' % (b['ms_per_microbatch_256'], b['tflops'], b['value'], b['e2e']['value']), b['clocks'], b['roofline']['frac'], b['parity']['ok'])
PY
