mkdir -p gpurun_out
B200CB_LIB=megengine_b200/lib/libb200cb_dbg.so timeout 300 python tools/dbg_timing.py bf16 pw > gpurun_out/dbg_pw.log 2>&1; cat gpurun_out/dbg_pw.log
