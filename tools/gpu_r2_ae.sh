O=gpurun_out/r2ae
mkdir -p $O
( time timeout 900 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err ) 2>&1 | grep real; tail -2 $O/bench.err
python -c "
import json; b=json.load(open('$O/bench.json')); print(json.dumps(b['extra']['config1_resnet18_fp32_nchw'], indent=1)); print(b['ms_per_microbatch_256'], b['clocks'])"
