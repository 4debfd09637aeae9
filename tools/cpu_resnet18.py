#!/usr/bin/env python
"""BASELINE.json configs[0] on the host cores: the ResNet-18 conv stack (fp32 NCHW, N = 8, 224x224,
bias + ReLU) through the reference's own x86 / fallback megdnn ConvBias (oracle/_ref), 1 thread and all
threads, median of >= 5 runs after 2 warm-ups.  Test infrastructure / baseline
only -- nothing here touches the product.

    python tools/cpu_resnet18.py --out profiles/r01/cpu_resnet18_x86.json
"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import _cases as K, _oracle as O  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--runs", type=int, default=5)
ap.add_argument("--out", default="")
a = ap.parse_args()
assert O.have_ref(), "oracle/_ref not built (oracle/build_ref.sh)"
cores = len(os.sched_getaffinity(0))
res = {"config": "ResNet-18 conv stack, fp32 NCHW, batch %d, 224x224, bias+ReLU, megdnn x86 handle" % a.batch,
       "cores": cores, "threads": {}}
for threads in sorted({1, cores}):
    rows, tot_s, tot_f = [], 0.0, 0.0
    for layer in K.RESNET18_LAYERS:
        d = K.resnet18_desc(layer, a.batch)
        src, flt, bias, _ = K.gen_inputs(d, 1)
        _dst, secs, algo = O.ref_convbias(d, src, flt, bias, None, handle_kind=0, threads=threads, repeat=a.runs + 2)
        t = float(np.median(list(secs)[2:]))
        fl = 2.0 * d.n * d.oc * d.oh * d.ow * d.ic * d.fh * d.fw
        rows.append({"layer": layer[0], "count": layer[-1], "algo": algo, "ms": round(t * 1e3, 3),
                     "gflops": round(fl / t / 1e9, 1)})
        tot_s += t * layer[-1]
        tot_f += fl * layer[-1]
        print("%2d thr %-9s x%d %-40s %8.3f ms %7.1f GFLOP/s" % (threads, layer[0], layer[-1], algo[:40], t * 1e3, fl / t / 1e9), flush=True)
    res["threads"][str(threads)] = {"stack_ms": round(tot_s * 1e3, 2), "images_per_s": round(a.batch / tot_s, 2),
                                    "gflops": round(tot_f / tot_s / 1e9, 1), "layers": rows}
    print("%2d threads: stack %.1f ms, %.1f images/s, %.1f GFLOP/s" % (threads, tot_s * 1e3, a.batch / tot_s, tot_f / tot_s / 1e9))
if a.out:
    json.dump(res, open(a.out, "w"), indent=1)
