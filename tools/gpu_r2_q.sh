O=gpurun_out/r2q
mkdir -p $O
for H in wc pinned wc pinned; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 6 --warmup 3 --no-validate --host-input $H > $O/b_$H.json 2> $O/b_$H.err
python - <<PY
import json
b=json.load(open('$O/b_$H.json')); print('$H', 'value %.0f'%b['value'], 'e2e %.0f'%b['e2e']['value'], 'GB/s %.1f'%b['e2e']['host_copy_gbs_all_gpus'], b['e2e']['host_input_memory'])
PY
done
