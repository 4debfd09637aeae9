O=gpurun_out/r2bn
mkdir -p $O
B="python bench.py --global-batch 256 --steps 2 --warmup 3 --no-validate --no-extras --no-cpu-baseline"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:'igemm|fold' -s 162 -c 54 --csv --log-file $O/launches_bf16.csv $B > /tmp/l1.log 2>&1; echo launches_bf16=$?
python tools/launches_to_traffic.py $O/launches_bf16.csv bf16 "ncu --metrics $M --clock-control none -k regex:igemm|fold -s 162 -c 54 $B (profiles/r02/launches_bf16.csv): the 54 launches of one 256-image micro-batch" > $O/dram_traffic_bf16.json
