set -x
mkdir -p gpurun_out
timeout 120 python tools/dw_one.py 58 28 28 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw3x3 -s 4 -c 1 -f -o gpurun_out/prof_dwbulk python tools/dw_one.py 58 28 28 1 bf16 4 > gpurun_out/ncu_dwbulk.log 2>&1; echo ncu=$?
