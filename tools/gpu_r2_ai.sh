O=gpurun_out/r2ai
mkdir -p $O
for rep in 1 2; do for L in libb200cb libb200cb_backoff300 libb200cb_backoff1000; do for K in bf16 qs8; do
B200CB_LIB=megengine_b200/lib/$L.so timeout 300 python bench.py --dtype $K --global-batch 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-validate > $O/b_${K}_${L}_$rep.json 2> $O/b_${K}_${L}_$rep.err
python - <<PY
import json
b=json.load(open('$O/b_${K}_${L}_$rep.json')); print('$K $L rep$rep', 'mb_ms %.4f'%b['ms_per_microbatch_256'], b['clocks']['sm_mhz'])
PY
done; done; done
for L in libb200cb libb200cb_backoff300 libb200cb_backoff1000; do
B200CB_LIB=megengine_b200/lib/$L.so timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras --no-validate > $O/s_${L}.json 2> $O/s_${L}.err
python - <<PY
import json
b=json.load(open('$O/s_${L}.json')); print('sustained bf16 $L', 'mb_ms %.4f'%b['ms_per_microbatch_256'], b['clocks']['sm_mhz'])
PY
done
B200CB_LIB=megengine_b200/lib/libb200cb_backoff300.so timeout 600 python -m pytest tests -m gpu -x -q -k "resnet50 or full_size or every_algo or dw or nchw4" 2>&1 | tail -2
