set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "depthwise or dw or dropin" > gpurun_out/pytest_dw.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_dw.log
timeout 120 python tools/bench_layers.py --set dw --dtype qs8 --out gpurun_out/dw_qs8.json > gpurun_out/dw_qs8.log 2>&1; tail -1 gpurun_out/dw_qs8.log
