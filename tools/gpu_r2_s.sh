O=gpurun_out/r2s
mkdir -p $O
B200CB_LIB=megengine_b200/lib/libb200cb_dbg.so timeout 300 python tools/dbg_timing.py bf16 > $O/dbg_bf16.log 2>&1; cut -c1-330 $O/dbg_bf16.log
B200CB_LIB=megengine_b200/lib/libb200cb_dbg.so timeout 300 python tools/dbg_timing.py qs8 > $O/dbg_qs8.log 2>&1
for NZ in 4 8; do for ST in 2 3; do
B200CB_LIB=megengine_b200/lib/libb200cb_tools.so B200CB_NZ=$NZ B200CB_STAGES=$ST timeout 120 python tools/sweep_resnet.py --dtype bf16 --only .c3 > $O/sweep_c3_nz${NZ}_st$ST.txt 2>&1; echo "nz $NZ stages<=$ST"; grep "^l" $O/sweep_c3_nz${NZ}_st$ST.txt | cut -c1-260
done; done
