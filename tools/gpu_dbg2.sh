set -x
mkdir -p gpurun_out
B200CB_LIB=megengine_b200/lib/libb200cb_dbg.so timeout 300 python tools/dbg_timing.py bf16 > gpurun_out/dbg_bf16.log 2>&1; cat gpurun_out/dbg_bf16.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -s 159 -c 3 -f -o gpurun_out/prof_conv1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv1.log 2>&1; echo ncu=$?
