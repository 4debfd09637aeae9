set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -x -q -k "depthwise or nchw4 or rows or stem or int8_epilogue or every_algo" > gpurun_out/sanitize_memcheck.log 2>&1; echo memcheck=$?; tail -8 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -x -q -k "dw3_tma_c144_56 or dw3s1_bf16" > gpurun_out/sanitize_racecheck.log 2>&1; echo racecheck=$?; tail -6 gpurun_out/sanitize_racecheck.log
