O=gpurun_out/r2az
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "selftest or int8 or qs8 or nchw4 or resnet50_layers_n1 or golden or resident or depthwise or nn_gpu or graph" > $O/pytest_int8.log 2>&1; echo pytest=$?; tail -3 $O/pytest_int8.log
for rep in 1 2; do
for L in prev new; do
LIB=megengine_b200/lib/libb200cb_prev.so; [ $L = new ] && LIB=megengine_b200/lib/libb200cb.so
B200CB_LIB=$LIB timeout 300 python bench.py --dtype qs8 --global-batch 2048 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --layer-report $O/layers_${L}_$rep.json > $O/bench_${L}_$rep.json 2> $O/err_${L}_$rep.txt
python -c "
import json; b=json.load(open('$O/bench_${L}_$rep.json')); print('$L $rep ms/mb %.4f' % b['ms_per_microbatch_256'], b['clocks'], b['parity']['ok'])"
done; done
