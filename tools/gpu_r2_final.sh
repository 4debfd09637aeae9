# round-2 final artefacts: smoke, full parity suite, driver-style bench lines, reference arm, ncu launch lists with DRAM
# bytes (the kernels' share of the step), full captures of the new kernels (int8 rows-mode l1.c2, pooling)
O=gpurun_out/r2final2
T=/tmp/ncu_r2f2
mkdir -p $O $T
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo smoke=$?; tail -4 $O/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 --layer-report $O/layers_bf16.json > $O/bench_bf16.json 2> $O/bench_bf16.err; echo bench=$?; tail -2 $O/bench_bf16.err
timeout 600 python bench.py --dtype qs8 --steps 5 --warmup 3 --no-cpu-baseline --no-extras --layer-report $O/layers_qs8.json > $O/bench_qs8.json 2> $O/bench_qs8.err; echo bench_qs8=$?
timeout 300 python bench.py --global-batch 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-extras --no-validate > $O/bench_bf16_short.json 2> $O/bench_bf16_short.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo ref=$?
B="python bench.py --global-batch 256 --steps 2 --warmup 3 --no-validate --no-extras --no-cpu-baseline"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:'igemm|fold' -s 162 -c 54 --csv --log-file $O/launches_bf16.csv $B > $T/l1.log 2>&1; echo launches_bf16=$?
timeout 900 ncu --metrics $M --clock-control none -k regex:'igemm|fold' -s 162 -c 54 --csv --log-file $O/launches_qs8.csv $B --dtype qs8 > $T/l2.log 2>&1; echo launches_qs8=$?
python tools/launches_to_traffic.py $O/launches_bf16.csv bf16 "ncu --metrics $M --clock-control none -k regex:igemm|fold -s 162 -c 54 $B (profiles/r02/launches_bf16.csv): the 54 launches of one 256-image micro-batch" > $O/dram_traffic_bf16.json
python tools/launches_to_traffic.py $O/launches_qs8.csv qs8 "ncu --metrics $M --clock-control none -k regex:igemm|fold -s 162 -c 54 $B --dtype qs8 (profiles/r02/launches_qs8.csv): the 54 launches of one 256-image micro-batch" > $O/dram_traffic_qs8.json
# full captures: int8 l1.c2 in rows mode with 64-byte rows (launch 2 of the qs8 step), the stem max pool
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -s 164 -c 1 -f -o $T/prof_qs8_l1c2 $B --dtype qs8 > $T/f1.log 2>&1; echo full_qs8=$?
python tools/ncu_summary.py $T/prof_qs8_l1c2.ncu-rep "l1.c2 QuantizedS8 (rows mode, 64-byte rows)" > $O/ncu_igemm_qs8_l1c2_rows64_summary.txt 2>&1
python tools/ncu_stalls.py $T/prof_qs8_l1c2.ncu-rep 0 > $O/ncu_stalls_qs8_l1c2_rows64.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pool_kernel -s 2 -c 1 -f -o $T/prof_pool python tools/bench_nn.py > $T/f2.log 2>&1; echo full_pool=$?
python tools/ncu_summary.py $T/prof_pool.ncu-rep "max pool 3x3/2 112x112x64 bf16 NHWC, N=256" > $O/ncu_pool_max_bf16_summary.txt 2>&1
python tools/ncu_stalls.py $T/prof_pool.ncu-rep 0 > $O/ncu_stalls_pool_max_bf16.txt 2>&1
rm -f $T/*.ncu-rep
python - <<'PY'
import json
for f in ['bf16','qs8','bf16_short']:
    try:
        b=json.load(open('gpurun_out/r2final2/bench_%s.json'%f)); print(f, 'ms/mb %.4f TF %.1f value %.0f e2e %.0f' % (b['ms_per_microbatch_256'], b['tflops'], b['value'], b['e2e']['value']), b['clocks'], b['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
print(open('gpurun_out/r2final2/bench_reference.json').read()[:600])
PY
du -sh $O; ls $O
