# GPU session script (run under gpurun): parity tests, bench (bf16 + int8), ncu launch list + full capture
set -x
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke=$?; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo pytest=$?; tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --layer-report gpurun_out/layers_bf16.json > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo bench=$?; cat gpurun_out/bench_bf16.json; tail -3 gpurun_out/bench_bf16.err
timeout 600 python bench.py --steps 20 --warmup 3 --dtype qs8 --no-cpu-baseline --layer-report gpurun_out/layers_qs8.json > gpurun_out/bench_qs8.json 2> gpurun_out/bench_qs8.err; echo benchq=$?; cat gpurun_out/bench_qs8.json; tail -3 gpurun_out/bench_qs8.err
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'igemm|fold|dwconv' -s 162 -c 54 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo ncu1=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -s 159 -c ${NCU_COUNT:-12} -f -o gpurun_out/prof_igemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo ncu2=$?
fi
ls -la gpurun_out
