O=gpurun_out/r2t
mkdir -p $O
export B200CB_LIB=megengine_b200/lib/libb200cb_tools.so
timeout 200 python tools/exp_cluster.py l3.bN.c2 l3.bN.c1 l4.bN.c1 l2.bN.c2 > $O/exp_time.txt 2>&1; cat $O/exp_time.txt
M=gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,sm__cycles_elapsed.max
EXP_REPS=1 timeout 600 ncu --metrics $M --clock-control none -k regex:igemm --csv --log-file $O/exp_ncu.csv python tools/exp_cluster.py l3.bN.c2 l3.bN.c1 > $O/exp_ncu.log 2>&1; echo ncu=$?
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r2t/exp_ncu.csv') if l.startswith('"')))
hdr = rows[0]; iid = hdr.index('ID'); im = hdr.index('Metric Name'); iv = hdr.index('Metric Value'); ik = hdr.index('Kernel Name')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[iid], {'k': r[ik][-40:]})[r[im]] = r[iv]
for i, v in d.items():
    print(i, v['k'], ' '.join('%s=%s' % (k.split('.')[0][-28:], v[k]) for k in v if k != 'k'))
PY
