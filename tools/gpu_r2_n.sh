O=gpurun_out/r2n
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -6 $O/pytest_gpu.log | cut -c1-300
for K in qs8 f32 bf16; do
timeout 300 python tools/bench_layers.py --set dw --dtype $K --out $O/dw_$K.json > $O/dw_$K.log 2>&1; tail -1 $O/dw_$K.log | cut -c1-400
done
