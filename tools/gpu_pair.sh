set -x
mkdir -p gpurun_out
B200CB_PAIR=1 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_pair.log 2>&1; echo pytest_pair=$?; tail -6 gpurun_out/pytest_pair.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep_resnet.py --dtype bf16 --out gpurun_out/sweep_bf16.json > gpurun_out/sweep_bf16.log 2>&1; echo sweep=$?; tail -30 gpurun_out/sweep_bf16.log
timeout 600 python tools/sweep_resnet.py --dtype qs8 --out gpurun_out/sweep_qs8.json > gpurun_out/sweep_qs8.log 2>&1; echo sweepq=$?; tail -30 gpurun_out/sweep_qs8.log
