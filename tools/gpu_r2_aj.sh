O=gpurun_out/r2aj
mkdir -p $O
timeout 600 python -m pytest tests/test_nn_gpu.py -q -x -k typecvt 2>&1 | tail -2
timeout 200 python tools/bench_nn.py 2>&1 | grep TypeCvt
( time timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err ) 2>&1 | grep real; tail -2 $O/bench.err
python -c "
import json; b=json.load(open('$O/bench.json')); print(b['extra']['latency_small_batch_bf16'])"
