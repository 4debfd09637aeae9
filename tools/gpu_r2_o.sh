O=gpurun_out/r2o
mkdir -p $O
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "depthwise or dw" > $O/pytest_dw.log 2>&1; echo pytest=$?; tail -3 $O/pytest_dw.log | cut -c1-300
timeout 300 python tools/bench_layers.py --set dw --dtype qs8 --out $O/dw_qs8.json > $O/dw_qs8.log 2>&1; grep -o '"ms": [0-9.]*, "gbs": [0-9.]*, "frac_hbm": [0-9.]*' $O/dw_qs8.log | tr '\n' ';'; tail -1 $O/dw_qs8.log | cut -c1-200
