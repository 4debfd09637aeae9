set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -x -q -k "depthwise or dw or nchw4 or rows or stem or int8_epilogue or every_algo" > gpurun_out/sanitize_memcheck.log 2>&1; echo memcheck=$?; tail -8 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -x -q -k "dw3_tma_c144_56 or dw3s1_bf16 or dw3_bulk_c58 or dw3_bulk_c36 or dw3_tma_7x7_n3 or dw3_tma_14s2_n5" > gpurun_out/sanitize_racecheck.log 2>&1; echo racecheck=$?; tail -6 gpurun_out/sanitize_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -x -q -k "dw3_bulk or dw3_tma" > gpurun_out/sanitize_synccheck.log 2>&1; echo synccheck=$?; tail -6 gpurun_out/sanitize_synccheck.log
