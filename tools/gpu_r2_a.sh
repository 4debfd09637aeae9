# round 2, call A: parity with the direct-store epilogue, A/B of B200CB_DIRECT (tools build), phase timing
set -x
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -5 $O/pytest_gpu.log
# every case again with the direct-store epilogue forced on / off
B200CB_LIB=megengine_b200/lib/libb200cb_tools.so B200CB_DIRECT=1 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $O/pytest_direct1.log 2>&1; echo pytest_d1=$?; tail -3 $O/pytest_direct1.log
B200CB_LIB=megengine_b200/lib/libb200cb_tools.so B200CB_DIRECT=0 timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > $O/pytest_direct0.log 2>&1; echo pytest_d0=$?; tail -3 $O/pytest_direct0.log
for D in 0 1; do
  for K in bf16 qs8; do
    B200CB_LIB=megengine_b200/lib/libb200cb_tools.so B200CB_DIRECT=$D timeout 300 python bench.py --steps 20 --warmup 3 --dtype $K --no-cpu-baseline --layer-report $O/layers_${K}_d$D.json > $O/bench_${K}_d$D.json 2> $O/bench_${K}_d$D.err; echo bench_${K}_d$D=$?
  done
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --layer-report $O/layers_bf16_prod.json > $O/bench_bf16_prod.json 2> $O/bench_bf16_prod.err; echo bench_prod=$?
for D in 0 1; do
  B200CB_LIB=megengine_b200/lib/libb200cb_dbg.so B200CB_DIRECT=$D timeout 300 python tools/dbg_timing.py bf16 > $O/dbg_bf16_d$D.log 2>&1
  B200CB_LIB=megengine_b200/lib/libb200cb_dbg.so B200CB_DIRECT=$D timeout 300 python tools/dbg_timing.py bf16 pw > $O/dbg_pw_d$D.log 2>&1
  B200CB_LIB=megengine_b200/lib/libb200cb_dbg.so B200CB_DIRECT=$D timeout 300 python tools/dbg_timing.py qs8 > $O/dbg_qs8_d$D.log 2>&1
done
for D in 0 1; do
  B200CB_LIB=megengine_b200/lib/libb200cb_tools.so B200CB_DIRECT=$D timeout 300 python tools/bench_layers.py --set pw --dtype bf16 --out $O/pw_bf16_d$D.json > $O/pw_bf16_d$D.log 2>&1
  B200CB_LIB=megengine_b200/lib/libb200cb_tools.so B200CB_DIRECT=$D timeout 300 python tools/bench_layers.py --set pw --dtype qs8 --out $O/pw_qs8_d$D.json > $O/pw_qs8_d$D.log 2>&1
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a/bench_*.json')):
    try:
        b=json.load(open(f)); print(f, 'ms_per_step %.4f'%b['ms_per_step'], 'TF %.1f'%b['tflops'], 'e2e %.0f'%b['e2e']['value'], b['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
