O=gpurun_out/r2y
mkdir -p $O
timeout 600 python -m pytest tests/test_nn_gpu.py -q -x 2>&1 | tail -30 > $O/pytest_nn.log; cat $O/pytest_nn.log
