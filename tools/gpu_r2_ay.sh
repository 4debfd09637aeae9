# instruction profiles (ncu --set full, per-SASS-line executed counts) of hot ResNet layers
O=gpurun_out/r2ay
T=/tmp/ncu_r2ay
mkdir -p $O $T
for spec in "l1.c3 qs8" "l1.ds qs8" "l1.c3 bf16" "l3.bN.c2 bf16" "l1.c2 bf16"; do
set -- $spec
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm -s 2 -c 1 -f -o $T/p python tools/one_layer_loop.py $1 $2 4 4 > $T/log.txt 2>&1; echo "$1 $2 rc=$?"
ncu -i $T/p.ncu-rep --page source --csv --print-source sass > $O/sass_$1_$2.csv 2>/dev/null
python tools/ncu_summary.py $T/p.ncu-rep "$1 $2" > $O/summary_$1_$2.txt 2>&1
done
ls -la $O
