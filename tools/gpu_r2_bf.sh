O=gpurun_out/r2bf
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 $O/pytest_gpu.log
for dt in bf16 f16 qs8; do timeout 300 python tools/bench_layers.py --set pw --dtype $dt --out $O/pw_$dt.json > $O/pw_$dt.log 2>&1; tail -1 $O/pw_$dt.log; done
python - <<'PY'
import json
for dt in ('bf16','qs8'):
    r=json.load(open('gpurun_out/r2bf/pw_%s.json'%dt))
    for y in r['rows']: print('   %-6s %-40s %s %.4f (%.2f)'%(dt, y['desc'].split(' d1')[0].split('n256 ')[1], y['algo'][-7:], y['ms'], y['frac_hbm']))
PY
