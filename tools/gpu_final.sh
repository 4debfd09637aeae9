# final artefacts of the round: smoke, parity, bench (bf16 / f16 / qs8), dw / pw tables, ncu launch list + full captures
set -x
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke=$?; tail -5 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --layer-report gpurun_out/layers_bf16.json > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo bench=$?
timeout 600 python bench.py --dtype qs8 --no-cpu-baseline --layer-report gpurun_out/layers_qs8.json > gpurun_out/bench_qs8.json 2> gpurun_out/bench_qs8.err
timeout 600 python bench.py --dtype f16 --no-cpu-baseline --layer-report gpurun_out/layers_f16.json > gpurun_out/bench_f16.json 2> gpurun_out/bench_f16.err
timeout 600 python bench.py --conv1-pixel-bytes 16 --no-cpu-baseline > gpurun_out/bench_bf16_ic8.json 2> gpurun_out/bench_bf16_ic8.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
for dt in bf16 qs8 f16 f32; do timeout 120 python tools/bench_layers.py --set dw --dtype $dt --out gpurun_out/dw_$dt.json > gpurun_out/dw_$dt.log 2>&1; tail -1 gpurun_out/dw_$dt.log; done
for dt in bf16 qs8; do timeout 120 python tools/bench_layers.py --set pw --dtype $dt --out gpurun_out/pw_$dt.json > gpurun_out/pw_$dt.log 2>&1; tail -1 gpurun_out/pw_$dt.log; done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'igemm|fold|dwconv|transpose' -s 159 -c 53 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo ncu1=$?
# full captures: conv1, l1.b0.c1, l1.c2, l1.ds, l1.c3 (launches 0..4 of the step) and l3.bN.c1/c2/c3 (33..35)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -s 159 -c 5 -f -o gpurun_out/prof_l1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1; echo ncu2=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -s 192 -c 3 -f -o gpurun_out/prof_l3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l3.log 2>&1; echo ncu3=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw3x3 -s 4 -c 2 -f -o gpurun_out/prof_dw3 python tools/dw_one.py 96 112 112 2 bf16 4 > gpurun_out/ncu_dw3.log 2>&1; echo ncu4=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw3x3 -s 4 -c 1 -f -o gpurun_out/prof_dwbulk python tools/dw_one.py 58 56 56 2 bf16 4 > gpurun_out/ncu_dwbulk.log 2>&1; echo ncu5=$?
# the reports are too large to travel back (gpurun_out/ is capped at 64 MiB): condense them here
python tools/ncu_summary.py gpurun_out/prof_l1.ncu-rep conv1 l1.b0.c1 l1.c2 l1.ds l1.c3 > gpurun_out/ncu_igemm_l1_summary.txt 2>&1
python tools/ncu_summary.py gpurun_out/prof_l3.ncu-rep l3.bN.c1 l3.bN.c2 l3.bN.c3 > gpurun_out/ncu_igemm_l3_pair_summary.txt 2>&1
python tools/ncu_summary.py gpurun_out/prof_dw3.ncu-rep "dw 96x112x112 s2 bf16" "dw 96x112x112 s2 bf16" > gpurun_out/ncu_dw3x3_summary.txt 2>&1
python tools/ncu_summary.py gpurun_out/prof_dwbulk.ncu-rep "dw 58x56x56 s2 bf16 (bulk rows)" >> gpurun_out/ncu_dw3x3_summary.txt 2>&1
rm -f gpurun_out/prof_l1.ncu-rep gpurun_out/prof_l3.ncu-rep gpurun_out/prof_dw3.ncu-rep
du -sh gpurun_out
python - <<'PY'
import json
for f in ['bf16','qs8','f16','bf16_ic8']:
    try:
        b=json.load(open('gpurun_out/bench_%s.json'%f)); print(f, 'ms_per_step %.4f TF %.1f value %.0f e2e %.0f' % (b['ms_per_step'], b['tflops'], b['value'], b['e2e']['value']), b['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
