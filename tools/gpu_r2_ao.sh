O=gpurun_out/r2ao
mkdir -p $O
export B200CB_LIB=megengine_b200/lib/libb200cb_tools.so
for dt in bf16 f16; do
for a in -1 1 2 3; do
B200CB_BRES=0 timeout 300 python tools/bench_layers.py --set pw --dtype $dt --algo $a --out $O/pw_${dt}_stream_a$a.json > /dev/null 2>&1
B200CB_BRES=1 timeout 300 python tools/bench_layers.py --set pw --dtype $dt --algo $a --out $O/pw_${dt}_res_a$a.json > /dev/null 2>&1
done; done
python - <<'PY'
import json,glob
for dt in ('bf16','f16'):
  for a in (-1,1,2,3):
    try:
        s=json.load(open('gpurun_out/r2ao/pw_%s_stream_a%d.json'%(dt,a))); r=json.load(open('gpurun_out/r2ao/pw_%s_res_a%d.json'%(dt,a)))
    except Exception as e:
        print(dt,a,'ERR',e); continue
    print(dt,'algo',a)
    for x,y in zip(s['layers'],r['layers']):
        if 'error' in x or 'error' in y: print('   ', x.get('desc','')[20:60], 'n/a'); continue
        print('   %s %s %.4f (%.2f) -> %s %.4f (%.2f)'%(x['desc'][21:58], x['algo'][-7:], x['ms'], x['frac_hbm'], y['algo'][-7:], y['ms'], y['frac_hbm']))
PY
