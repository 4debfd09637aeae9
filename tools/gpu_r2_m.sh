O=gpurun_out/r2m
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -25 $O/pytest_gpu.log | cut -c1-300
