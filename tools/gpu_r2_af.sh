O=gpurun_out/r2af
T=/tmp/ncu_r2af
mkdir -p $O $T
B="python bench.py --global-batch 256 --steps 2 --warmup 3 --no-validate --no-extras --no-cpu-baseline --dtype qs8"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -s 165 -c 1 -f -o $T/prof_qs8_l1ds $B > $T/f1.log 2>&1; echo full=$?
python tools/ncu_summary.py $T/prof_qs8_l1ds.ncu-rep "l1.ds QuantizedS8 (1x1 64->256, epilogue bound)" > $O/ncu_igemm_qs8_l1ds_summary.txt 2>&1
python tools/ncu_stalls.py $T/prof_qs8_l1ds.ncu-rep 0 > $O/ncu_stalls_qs8_l1ds.txt 2>&1
python tools/ncu_hot.py $T/prof_qs8_l1ds.ncu-rep 0 60 > $O/ncu_hot_qs8_l1ds.txt 2>&1
ncu -i $T/prof_qs8_l1ds.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2] if len(rows)>2 else rows[1]
for k in h:
    if any(t in k for t in ('inst_executed_pipe','issue_active','pipe_fma','pipe_alu','pipe_xu','pipe_lsu','smsp__inst_executed.avg.per_cycle','smsp__issue','warp_issue_stalled')):
        print(k, r[h.index(k)])
" > $O/ncu_pipes_qs8_l1ds.txt 2>&1
cat $O/ncu_igemm_qs8_l1ds_summary.txt | head -20; cat $O/ncu_stalls_qs8_l1ds.txt | cut -c1-700; head -45 $O/ncu_hot_qs8_l1ds.txt; head -50 $O/ncu_pipes_qs8_l1ds.txt
