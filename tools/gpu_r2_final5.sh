# round-2 final artefacts: smoke, full parity suite, driver-style bench lines, reference arm, ncu launch lists with DRAM
# bytes (the kernels' share of the step), full captures of the new kernels (int8 rows-mode l1.c2, pooling)
O=gpurun_out/r2final5
T=/tmp/ncu_r2f5
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
# full capture: the low-K pointwise layer 16->96 112x112 bf16 with the resident filter tile (tools/pw_one.py)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -s 3 -c 1 -f -o $T/prof_pw python tools/pw_one.py 16 96 112 112 bf16 5 > $T/f1.log 2>&1; echo full_pw=$?; tail -3 $T/f1.log
python tools/ncu_summary.py $T/prof_pw.ncu-rep "pointwise 16->96 112x112 bf16 NHWC, N=256, resident filter tile (b_res)" > $O/ncu_igemm_bf16_pw16to96_bres_summary.txt 2>&1
python tools/ncu_stalls.py $T/prof_pw.ncu-rep 0 > $O/ncu_stalls_bf16_pw16to96_bres.txt 2>&1
for dt in bf16 qs8 f32; do timeout 300 python tools/bench_layers.py --set dw --dtype $dt --out $O/dw_$dt.json > /dev/null 2>&1; done
for dt in bf16 f16 qs8; do timeout 300 python tools/bench_layers.py --set pw --dtype $dt --out $O/pw_$dt.json > /dev/null 2>&1; done
timeout 300 python tools/bench_nn.py --out $O/nn_ops.json > $O/nn_ops.log 2>&1; tail -3 $O/nn_ops.log
rm -f $T/*.ncu-rep
python - <<'PY'
import json
for f in ['bf16','qs8','bf16_short']:
    try:
        b=json.load(open('gpurun_out/r2final5/bench_%s.json'%f)); print(f, 'ms/mb %.4f TF %.1f value %.0f e2e %.0f' % (b['ms_per_microbatch_256'], b['tflops'], b['value'], b['e2e']['value']), b['clocks'], b['roofline']['frac'])
    except Exception as e: print(f, 'ERR', e)
print(open('gpurun_out/r2final5/bench_reference.json').read()[:600])
PY
du -sh $O; ls $O
