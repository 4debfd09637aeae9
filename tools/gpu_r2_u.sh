O=gpurun_out/r2u
mkdir -p $O
export B200CB_LIB=megengine_b200/lib/libb200cb_tools.so
timeout 300 python tools/exp_c3.py l3.c3 l4.c3 l2.c3 l1.c3 > $O/exp_c3.txt 2>&1; cat $O/exp_c3.txt
