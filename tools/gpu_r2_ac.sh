O=gpurun_out/r2ac
mkdir -p $O
for rep in 1 2; do for L in libb200cb libb200cb_hint1000 libb200cb_hint20000; do
B200CB_LIB=megengine_b200/lib/$L.so timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extras --no-validate > $O/b_${L}_$rep.json 2> $O/b_${L}_$rep.err
python - <<PY
import json
b=json.load(open('$O/b_${L}_$rep.json')); print('$L rep$rep', 'mb_ms %.4f'%b['ms_per_microbatch_256'], b['clocks'])
PY
done; done
