set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -4 gpurun_out/pytest_gpu.log
timeout 120 python tools/bench_layers.py --set pw --dtype bf16 --out gpurun_out/pw_bf16.json > gpurun_out/pw_bf16.log 2>&1; tail -1 gpurun_out/pw_bf16.log
timeout 120 python tools/bench_layers.py --set pw --dtype qs8 --out gpurun_out/pw_qs8.json > gpurun_out/pw_qs8.log 2>&1; tail -1 gpurun_out/pw_qs8.log
