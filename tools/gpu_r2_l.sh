O=gpurun_out/r2l
mkdir -p $O
export B200CB_LIB=megengine_b200/lib/libb200cb_tools.so
for rep in 1 2; do for R in 0 1; do for K in bf16 qs8; do
B200CB_RING=$R timeout 300 python bench.py --steps 3 --dtype $K --no-cpu-baseline --no-extras --no-validate --layer-report $O/layers_${K}_ring${R}_$rep.json > $O/bench_${K}_ring${R}_$rep.json 2> $O/bench_${K}_ring${R}_$rep.err
python - <<PY
import json
b=json.load(open('$O/bench_${K}_ring${R}_$rep.json')); L=json.load(open('$O/layers_${K}_ring${R}_$rep.json'))['layers']
print('$K ring$R rep$rep', 'mb_ms %.4f'%b['ms_per_microbatch_256'], b['clocks']['sm_mhz'], 'conv1 %.4f'%[r for r in L if r['name']=='conv1'][0]['ms'], 'sum %.3f'%sum(r['ms']*r['count'] for r in L))
PY
done; done; done
