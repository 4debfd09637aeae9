set -x
O=gpurun_out/r2c
mkdir -p $O
export B200CB_LIB=megengine_b200/lib/libb200cb_tools.so
for L in l1.c3 l2.c3 l1.ds; do
timeout 120 python tools/sweep_resnet.py --dtype bf16 --only $L --reps 20 --verbose > $O/sweep_$L.txt 2>&1; tail -4 $O/sweep_$L.txt
done
timeout 600 compute-sanitizer --tool memcheck python tools/sweep_resnet.py --dtype bf16 --only l1.c3 --batch 8 --reps 2 --verbose > $O/memcheck_l1c3.txt 2>&1; tail -30 $O/memcheck_l1c3.txt
