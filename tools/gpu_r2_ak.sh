O=gpurun_out/r2ak
T=/tmp/ncu_r2ak
mkdir -p $O $T
B="python bench.py --global-batch 256 --steps 2 --warmup 3 --no-validate --no-extras --no-cpu-baseline"
# bf16 step: launches 159.. ; l3 block: l3.bN.c1, l3.bN.c2, l3.c3 around index 187 (see tools/gpu_r2_ncu.sh)
timeout 600 ncu --set full --clock-control none -k regex:igemm -s 187 -c 3 -f -o $T/prof_l3 $B > $T/f1.log 2>&1; echo full=$?
python tools/ncu_top.py $T/prof_l3.ncu-rep 35 > $O/ncu_top_bf16_l3.txt 2>&1; cat $O/ncu_top_bf16_l3.txt | cut -c1-150
