O=gpurun_out/r2bc
mkdir -p $O
timeout 900 python -m pytest tests/test_nn_gpu.py -m gpu -x -q > $O/pytest_nn.log 2>&1; echo pytest_nn=$?; tail -4 $O/pytest_nn.log
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "selftest or int8 or qs8" > $O/pytest_int8.log 2>&1; echo pytest_int8=$?; tail -2 $O/pytest_int8.log
timeout 300 python tools/bench_nn.py --out $O/nn_ops.json 2>&1 | grep -i "add\|QFUSE"
