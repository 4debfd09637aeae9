# compute-sanitizer on the resident-filter-tile path (memcheck: every forced case; racecheck: a few multi-tile cases)
O=gpurun_out/r2bb
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -x -q -k "resident_filter_tile_forced" > $O/sanitize_memcheck_bres.txt 2>&1; echo memcheck=$?; tail -4 $O/sanitize_memcheck_bres.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_parity_gpu.py -x -q -k "resident_filter_tile_full_size and (l2.bN.c1 or 24x56x144)" > $O/sanitize_racecheck_bres.txt 2>&1; echo racecheck=$?; tail -4 $O/sanitize_racecheck_bres.txt
