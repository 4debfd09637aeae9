# depthwise regression on a B200: parity subset + the 17-layer table (bf16)
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "depthwise or dw or dropin" > gpurun_out/pytest_dw.log 2>&1; echo pytest=$?; tail -5 gpurun_out/pytest_dw.log
timeout 120 python tools/bench_layers.py --set dw --dtype bf16 --out gpurun_out/dw_bf16.json > gpurun_out/dw_bf16.log 2>&1; tail -1 gpurun_out/dw_bf16.log
