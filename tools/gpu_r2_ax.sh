O=gpurun_out/r2ax
mkdir -p $O
for rep in 1 2; do
for L in prev new; do
LIB=megengine_b200/lib/libb200cb_prev.so; [ $L = new ] && LIB=megengine_b200/lib/libb200cb.so
B200CB_LIB=$LIB timeout 300 python bench.py --global-batch 2048 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --no-validate --layer-report $O/layers_${L}_$rep.json > $O/bench_${L}_$rep.json 2> $O/err_${L}_$rep.txt
python -c "
import json; b=json.load(open('$O/bench_${L}_$rep.json')); print('$L $rep ms/mb %.4f' % b['ms_per_microbatch_256'], b['clocks'])"
done; done
