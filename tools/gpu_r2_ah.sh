O=gpurun_out/r2ah
T=/tmp/ncu_r2ah
mkdir -p $O $T
B="python bench.py --global-batch 256 --steps 2 --warmup 3 --no-validate --no-extras --no-cpu-baseline --dtype qs8"
timeout 600 ncu --set full --clock-control none -k regex:igemm -s 162 -c 6 -f -o $T/prof_qs8_l1 $B > $T/f1.log 2>&1; echo full=$?
python tools/ncu_summary.py $T/prof_qs8_l1.ncu-rep a b c d e f 2>/dev/null | grep -E "launch|Kernel Name|time_duration" > $O/order.txt; cat $O/order.txt
python tools/ncu_pipes.py $T/prof_qs8_l1.ncu-rep > $O/ncu_pipes_qs8_layer1.txt 2>&1; cat $O/ncu_pipes_qs8_layer1.txt | cut -c1-150
