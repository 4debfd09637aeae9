O=gpurun_out/r2bk
mkdir -p $O
timeout 120 ./tools/microbench/mma_shift > $O/microbench_mma_shift.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo pytest=$?; tail -2 $O/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo smoke=$?
