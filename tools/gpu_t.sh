set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -25 gpurun_out/pytest_gpu.log
