O=gpurun_out/r2r
mkdir -p $O
for rep in 1 2; do for L in tools hint2k hint20k; do for K in bf16 qs8; do
B200CB_LIB=megengine_b200/lib/libb200cb_$L.so timeout 300 python bench.py --global-batch 2048 --steps 3 --dtype $K --no-cpu-baseline --no-extras --no-validate > $O/b_${K}_${L}_$rep.json 2> $O/b_${K}_${L}_$rep.err
python - <<PY
import json
b=json.load(open('$O/b_${K}_${L}_$rep.json')); print('$K $L rep$rep', 'mb_ms %.4f'%b['ms_per_microbatch_256'], b['clocks']['sm_mhz'])
PY
done; done; done
for L in tools hint2k; do
B200CB_LIB=megengine_b200/lib/libb200cb_$L.so timeout 300 python tools/bench_layers.py --set pw --dtype bf16 > $O/pw_bf16_$L.log 2>&1; tail -1 $O/pw_bf16_$L.log | cut -c1-160
B200CB_LIB=megengine_b200/lib/libb200cb_$L.so timeout 300 python tools/bench_layers.py --set pw --dtype qs8 > $O/pw_qs8_$L.log 2>&1; tail -1 $O/pw_qs8_$L.log | cut -c1-160
done
