O=gpurun_out/r2bg
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py tests/test_nn_gpu.py -x -q -k "pixel_pair_fold or quantized_add_all_byte_pairs or quantized_elemwise or pooling" > $O/sanitize_memcheck_fold_qadd.txt 2>&1; echo memcheck=$?; tail -4 $O/sanitize_memcheck_fold_qadd.txt
