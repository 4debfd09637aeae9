O=gpurun_out/r2be
mkdir -p $O
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "pixel_pair_fold or resident_filter_tile_full_size" > $O/pytest_fold.log 2>&1; echo pytest_fold=$?; tail -5 $O/pytest_fold.log
export B200CB_LIB=megengine_b200/lib/libb200cb_tools.so
for dt in bf16 qs8; do
for fold in 0 1; do
echo "== $dt fold=$fold"
B200CB_PWFOLD=$fold timeout 120 python tools/pw_one.py 32 16 112 112 $dt 30 | cut -c1-230
B200CB_PWFOLD=$fold B200CB_BRES=1 timeout 120 python tools/pw_one.py 32 16 112 112 $dt 30 | cut -c1-230
done; done
