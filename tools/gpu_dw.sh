set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "depthwise or dw or smoke or dropin" > gpurun_out/pytest_dw.log 2>&1; echo pytest=$?; tail -5 gpurun_out/pytest_dw.log
for dt in bf16 f16; do timeout 60 python tools/bench_layers.py --set dw --dtype $dt --out gpurun_out/dw_$dt.json > gpurun_out/dw_$dt.log 2>&1; python - <<PY
import json
d=json.load(open('gpurun_out/dw_$dt.json'))
for l in d['rows']: print(l['desc'][14:62], l['ms'], l['gbs'], l['frac_hbm'])
print(d['summary'])
PY
done
