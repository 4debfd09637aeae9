O=gpurun_out/r2ab
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "int8 or qs8 or nchw4 or bit_exact or resnet50" > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 $O/pytest_gpu.log
for F in nhwc nchw4; do
timeout 300 python bench.py --dtype qs8 --format $F --global-batch 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-validate > $O/bench_$F.json 2> $O/bench_$F.err; tail -1 $O/bench_$F.err
python -c "
import json; b=json.load(open('$O/bench_$F.json')); print('$F ms/microbatch', b.get('ms_per_microbatch_256'), b['clocks'])"
done
