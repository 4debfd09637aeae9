# A/B of the 16-bit epilogue restructuring on the residual (z) layers: previous tools build vs current, same box, interleaved
O=gpurun_out/r2aw
mkdir -p $O
for rep in 1 2; do
for L in prev tools; do
echo "== $L rep $rep"
B200CB_LIB=megengine_b200/lib/libb200cb_$L.so timeout 200 python tools/exp_c3.py l1.c3 l2.c3 l3.c3 l4.c3 2>&1 | grep "base\|no_z "
done; done
