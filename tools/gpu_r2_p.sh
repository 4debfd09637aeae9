O=gpurun_out/r2p
mkdir -p $O
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; echo rc=$?; tail -5 $O/bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
b=json.load(open('$O/bench_${N}gpu.json'))
print({k:b[k] for k in ('value','n_gpus','ms_per_step','ms_per_microbatch_256','tflops','scaling')}, b['clocks'])
print('e2e', b['e2e']['value'], b['e2e']['host_copy_gbs_all_gpus'])
print('parity', json.dumps(b['parity'])[:900])
print('host', b['host'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/ref_${N}gpu.json 2> $O/ref_${N}gpu.err; echo ref_rc=$?; cut -c1-200 $O/ref_${N}gpu.json
nvidia-smi topo -m 2>/dev/null | head -14
