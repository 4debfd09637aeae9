O=gpurun_out/r2am
mkdir -p $O
for K in bf16 qs8 f16; do
timeout 600 python bench.py --dtype $K --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_$K.json 2> $O/bench_$K.err; echo "$K rc=$?"; tail -2 $O/bench_$K.err
python -c "
import json; b=json.load(open('$O/bench_$K.json')); print('$K value %.0f e2e %.0f' % (b['value'], b['e2e']['value']), {k:b['e2e'][k] for k in ('h2d_bytes_per_gpu_microbatch','d2h_bytes_per_gpu_microbatch','result_check_vs_oracle','host_copy_gbs_all_gpus')}, b['clocks']['sm_mhz'])"
done
