O=gpurun_out/r2w
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -5 $O/pytest_gpu.log
B200CB_LIB=megengine_b200/lib/libb200cb_tools.so timeout 200 python tools/sweep_resnet.py --dtype qs8 --only l1.c2 > $O/sweep_l1c2_qs8.txt 2>&1; cat $O/sweep_l1c2_qs8.txt
timeout 300 python bench.py --dtype qs8 --global-batch 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --layer-report $O/layers_qs8.json > $O/bench_qs8.json 2> $O/bench_qs8.err; echo bench=$?; tail -2 $O/bench_qs8.err; python -c "
import json; b=json.load(open('$O/bench_qs8.json')); print('qs8 ms/microbatch', b.get('ms_per_microbatch_256'), b['clocks'])
for r in json.load(open('$O/layers_qs8.json')):
    print('%-10s x%d %s %.4f floor %.4f' % (r['name'], r['count'], r['algo'][-8:], r['ms'], r['floor_ms']))
"
