# where do the executed instructions of the low-K pointwise layer go (per CUDA source line and per SASS line)?
O=gpurun_out/r2as
T=/tmp/ncu_r2as
mkdir -p $O $T
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -s 3 -c 1 -f -o $T/prof_pw python tools/pw_one.py 16 96 112 112 bf16 5 > $T/f1.log 2>&1; echo full_pw=$?
ncu -i $T/prof_pw.ncu-rep --page source --csv --print-source cuda > $O/pw_src_cuda.csv 2> $O/err1.txt
ncu -i $T/prof_pw.ncu-rep --page source --csv --print-source sass > $O/pw_src_sass.csv 2> $O/err2.txt
ls -la $O
