# call H: sub-ring z + shared bias + packed 16-bit epilogue: parity, stress, bench, timing, sweep
O=gpurun_out/r2h
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 $O/pytest_gpu.log
B200CB_LIB=megengine_b200/lib/libb200cb_tools.so B200CB_PAIR=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $O/pytest_pair.log 2>&1; echo pytest_pair=$?; tail -2 $O/pytest_pair.log
for K in f16 bf16 qs8; do
  timeout 200 python tools/stress_step.py $K 40 > $O/stress_$K.log 2>&1; echo "stress $K rc=$? $(tail -1 $O/stress_$K.log | cut -c1-100)"
done
for K in bf16 f16 qs8; do
  timeout 300 python bench.py --steps 20 --warmup 3 --dtype $K --no-cpu-baseline --layer-report $O/layers_${K}.json > $O/bench_${K}.json 2> $O/bench_${K}.err; echo bench_${K}=$?
done
export B200CB_LIB=megengine_b200/lib/libb200cb_tools.so
B200CB_DIRECT=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layer-report $O/layers_bf16_d0.json > $O/bench_bf16_d0.json 2> $O/bench_bf16_d0.err
B200CB_DIRECT=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layer-report $O/layers_bf16_d1.json > $O/bench_bf16_d1.json 2> $O/bench_bf16_d1.err
B200CB_LIB=megengine_b200/lib/libb200cb_dbg.so timeout 300 python tools/dbg_timing.py bf16 > $O/dbg_bf16.log 2>&1
B200CB_LIB=megengine_b200/lib/libb200cb_dbg.so timeout 300 python tools/dbg_timing.py qs8 > $O/dbg_qs8.log 2>&1
timeout 300 python tools/bench_layers.py --set pw --dtype bf16 --out $O/pw_bf16.json > $O/pw_bf16.log 2>&1
timeout 300 python tools/bench_layers.py --set pw --dtype qs8 --out $O/pw_qs8.json > $O/pw_qs8.log 2>&1
timeout 900 python tools/sweep_resnet.py --dtype bf16 --out $O/sweep_bf16.json > $O/sweep_bf16.txt 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h/bench_*.json')):
    try:
        b=json.load(open(f)); print(f, 'ms_per_step %.4f'%b['ms_per_step'], 'TF %.1f'%b['tflops'], 'e2e %.0f'%b['e2e']['value'], b['clocks']['sm_mhz'])
    except Exception as e: print(f, 'ERR', e)
PY
