O=gpurun_out/r2ba
T=/tmp/ncu_r2ba
mkdir -p $O $T
for spec in "144 56 56 1" "32 112 112 1"; do
set -- $spec
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw3x3 -s 3 -c 1 -f -o $T/p python tools/dw_one.py $1 $2 $3 $4 qs8 5 > $T/log.txt 2>&1; echo "$spec rc=$?"; tail -1 $T/log.txt
ncu -i $T/p.ncu-rep --page source --csv --print-source sass > $O/sass_dw_c$1_$2_s$4_qs8.csv 2>/dev/null
python tools/ncu_summary.py $T/p.ncu-rep "dw c$1 $2x$3 s$4 qs8" > $O/summary_dw_c$1_$2_s$4.txt 2>&1
done
