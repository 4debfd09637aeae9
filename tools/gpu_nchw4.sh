set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --dtype qs8 --format nchw4 --no-cpu-baseline --layer-report gpurun_out/layers_qs8_nchw4.json > gpurun_out/bench_qs8_nchw4.json 2> gpurun_out/bench_qs8_nchw4.err; echo rc=$?; tail -3 gpurun_out/bench_qs8_nchw4.err
python -c "
import json; b=json.load(open('gpurun_out/bench_qs8_nchw4.json')); print('qs8 nchw4', b['ms_per_step'], b['value'], b['e2e']['value'], b['gpu_launches'])"
