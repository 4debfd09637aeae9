O=gpurun_out/r2z
mkdir -p $O
timeout 600 python -m pytest tests/test_nn_gpu.py -q -x 2>&1 | tail -5
timeout 300 python tools/bench_nn.py --out $O/nn.json 2>&1 | tail -30
