O=gpurun_out/r2bi
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 $O/pytest_gpu.log
for dt in bf16 qs8; do
timeout 300 python bench.py --dtype $dt --global-batch 2048 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --layer-report $O/layers_$dt.json > $O/bench_$dt.json 2> $O/err_$dt.txt
python -c "
import json; b=json.load(open('$O/bench_$dt.json')); print('$dt ms/mb %.4f' % b['ms_per_microbatch_256'], b['clocks'], b['parity']['ok'])
for r in json.load(open('$O/layers_$dt.json'))['layers']:
    if r['name'] in ('conv1','l1.c2','l1.c3','l2.bN.c2'): print('  ', r['name'], r['algo'][-7:], r['ms'])"
done
