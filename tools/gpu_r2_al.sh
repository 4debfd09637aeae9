O=gpurun_out/r2al
mkdir -p $O
export B200CB_LIB=megengine_b200/lib/libb200cb_tools.so
for J in 0 1; do for NZ in 4 8; do
echo "zjit $J nz $NZ"
B200CB_ZJIT=$J B200CB_NZ=$NZ timeout 200 python tools/sweep_resnet.py --dtype bf16 --only .c3 2>&1 | grep "^l" | cut -c1-250
done; done
B200CB_ZJIT=1 B200CB_NZ=4 timeout 600 python -m pytest tests -m gpu -x -q -k "resnet50 or full_size or z or every_algo" 2>&1 | tail -2
