# round 2 profiles: launch list (time + DRAM bytes) of one 256-image micro-batch, full captures of the
# kernels DESIGN.md quotes, condensed ON THE BOX into text summaries (the reports themselves are too big to
# travel).  Numbers printed by runs under ncu are never bench values.
O=gpurun_out/r2ncu
T=/tmp/ncu_r2
mkdir -p $O $T
B="python bench.py --global-batch 256 --steps 2 --warmup 3 --no-validate --no-extras --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'igemm|fold' -s 162 -c 54 --csv --log-file $O/launches_bf16.csv $B > $T/l1.log 2>&1; echo launches_bf16=$?
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'igemm|fold' -s 162 -c 54 --csv --log-file $O/launches_qs8.csv $B --dtype qs8 > $T/l2.log 2>&1; echo launches_qs8=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -s 159 -c 12 -f -o $T/prof_bf16_l1 $B > $T/f1.log 2>&1; echo full_l1=$?
python tools/ncu_summary.py $T/prof_bf16_l1.ncu-rep conv1 l1.b0.c1 l1.c2 l1.ds l1.c3 l1.bN.c1 l1.c2 l1.c3 l1.bN.c1 l1.c2 l1.c3 l2.b0.c1 > $O/ncu_igemm_bf16_conv1_layer1_summary.txt
for i in 0 2 4; do python tools/ncu_stalls.py $T/prof_bf16_l1.ncu-rep $i > $O/ncu_stalls_bf16_l1_launch$i.txt 2>&1; python tools/ncu_hot.py $T/prof_bf16_l1.ncu-rep $i 20 > $O/ncu_hot_bf16_l1_launch$i.txt 2>&1; done
rm -f $T/prof_bf16_l1.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -s 187 -c 3 -f -o $T/prof_bf16_l3 $B > $T/f3.log 2>&1; echo full_l3=$?
python tools/ncu_summary.py $T/prof_bf16_l3.ncu-rep l3.bN.c1 l3.bN.c2 l3.c3 > $O/ncu_igemm_bf16_layer3_summary.txt
for i in 0 1 2; do python tools/ncu_stalls.py $T/prof_bf16_l3.ncu-rep $i > $O/ncu_stalls_bf16_l3_launch$i.txt 2>&1; done
rm -f $T/prof_bf16_l3.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -s 206 -c 3 -f -o $T/prof_bf16_l4 $B > $T/f4.log 2>&1; echo full_l4=$?
python tools/ncu_summary.py $T/prof_bf16_l4.ncu-rep l4.bN.c1 l4.bN.c2 l4.c3 > $O/ncu_igemm_bf16_layer4_summary.txt
rm -f $T/prof_bf16_l4.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw3x3 -s 4 -c 1 -f -o $T/prof_dw_qs8 python tools/dw_one.py 144 56 56 1 qs8 4 > $T/d1.log 2>&1; echo dw_qs8=$?
python tools/ncu_summary.py $T/prof_dw_qs8.ncu-rep "MobileNetV2 dw c144 56x56 s1 QuantizedS8" > $O/ncu_dw3x3_s8_tma_summary.txt; python tools/ncu_stalls.py $T/prof_dw_qs8.ncu-rep 0 > $O/ncu_stalls_dw3x3_s8_tma.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw3x3 -s 4 -c 1 -f -o $T/prof_dw_f32 python tools/dw_one.py 144 56 56 1 f32 4 > $T/d2.log 2>&1; echo dw_f32=$?
python tools/ncu_summary.py $T/prof_dw_f32.ncu-rep "MobileNetV2 dw c144 56x56 s1 Float32" > $O/ncu_dw3x3_f32_summary.txt
du -sh $O; ls $O
