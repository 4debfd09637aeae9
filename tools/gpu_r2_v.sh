O=gpurun_out/r2v
mkdir -p $O
export B200CB_LIB=megengine_b200/lib/libb200cb_tools.so
timeout 300 python tools/exp_stages.py l3.bN.c2 l3.bN.c1 l4.bN.c1 l4.bN.c2 l2.bN.c2 l2.bN.c1 l3.b0.c1 > $O/exp_stages.txt 2>&1; cat $O/exp_stages.txt
