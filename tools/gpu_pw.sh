mkdir -p gpurun_out
for a in -1 0 1 2; do timeout 120 python tools/bench_layers.py --set pw --dtype bf16 --algo $a --out gpurun_out/pw_bf16_a$a.json > gpurun_out/pw_a$a.log 2>&1; tail -1 gpurun_out/pw_a$a.log; done
