mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:igemm -s 4 -c 1 -f -o gpurun_out/prof_pw python tools/pw_one.py 32 16 112 112 bf16 4 > gpurun_out/ncu_pw.log 2>&1; echo ncu=$?; tail -3 gpurun_out/ncu_pw.log
