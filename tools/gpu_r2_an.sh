# resident filter tile: parity of the new tests, then the whole GPU suite, the qs8 step and the int8 pointwise table
O=gpurun_out/r2an
mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "resident_filter" > $O/pytest_bres.log 2>&1; echo pytest_bres=$?; tail -5 $O/pytest_bres.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 $O/pytest_gpu.log
timeout 600 python bench.py --dtype qs8 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --layer-report $O/layers_qs8.json > $O/bench_qs8.json 2> $O/bench_qs8.err; echo bench_qs8=$?; tail -2 $O/bench_qs8.err
timeout 300 python tools/bench_layers.py --set pw --dtype qs8 --out $O/pw_qs8.json > $O/pw_qs8.log 2>&1; tail -12 $O/pw_qs8.log
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2an/bench_qs8.json')); print('qs8 ms/mb %.4f value %.0f e2e %.0f' % (b['ms_per_microbatch_256'], b['value'], b['e2e']['value']), b['clocks'], b.get('validation'))
PY
