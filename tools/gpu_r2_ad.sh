O=gpurun_out/r2ad
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_nn_gpu.py tests/test_graph_gpu.py -x -q > $O/sanitize_memcheck_nn.log 2>&1; echo memcheck_nn=$?; tail -4 $O/sanitize_memcheck_nn.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -x -q -k "rows or q_rows or int8_epilogue" > $O/sanitize_memcheck_rows.log 2>&1; echo memcheck_rows=$?; tail -4 $O/sanitize_memcheck_rows.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -x -q -k "q_rows64_c3_7x7 or q_rows64_c5 or rows64_c3_f16" > $O/sanitize_racecheck_rows64.log 2>&1; echo racecheck_rows64=$?; tail -4 $O/sanitize_racecheck_rows64.log
