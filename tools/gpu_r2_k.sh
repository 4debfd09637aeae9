O=gpurun_out/r2k
mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "stem or conv1 or nchw" > $O/pytest_stem.log 2>&1; echo pytest=$?; tail -2 $O/pytest_stem.log
export B200CB_LIB=megengine_b200/lib/libb200cb_dbg.so
for R in 0 1; do for K in bf16 qs8; do
B200CB_RING=$R timeout 120 python tools/dbg_timing.py $K conv1 > $O/dbg_${K}_ring$R.log 2>&1; cat $O/dbg_${K}_ring$R.log | cut -c1-300
done; done
export B200CB_LIB=megengine_b200/lib/libb200cb_tools.so
for R in 0 1; do for K in bf16 qs8; do
B200CB_RING=$R timeout 120 python tools/sweep_resnet.py --dtype $K --only conv1 > $O/sweep_${K}_ring$R.log 2>&1; head -1 $O/sweep_${K}_ring$R.log | cut -c1-200
done; done
