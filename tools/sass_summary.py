#!/usr/bin/env python
"""SASS listing summary of the shipped library: which Blackwell instructions each kernel family contains
(UTCHMMA / UTCIMMA = tcgen05.mma f16 / i8, .2CTA = cta_group::2, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA
load / store, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, IDP.4A = dp4a, FFMA2 / FADD2 / FMUL2 = packed
fp32x2, HMNMX2 = packed 16-bit max).     python tools/sass_summary.py > profiles/rNN/sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "megengine_b200", "lib", "libb200cb.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
keys = ["UTCHMMA", "UTCIMMA", ".2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "IM2COL", "MULTICAST", "SYNCS",
        "FFMA2", "FADD2", "FMUL2", "HMNMX2", "IDP.4A", "PRMT", "UBLKCP", "ELECT"]
fam = collections.defaultdict(lambda: [0, collections.Counter(), 0])
tot = collections.Counter()
for f in funcs:
    name = f.split("\n", 1)[0]
    fn = name[:40]
    for cand in ("igemm_kernel", "dw3x3_s8_tma_kernel", "dw3x3_s8_kernel", "dw3x3_tma_kernel", "dw3x3_bulk_kernel",
                 "dw3x3_kernel", "dwconv_nhwc_kernel", "transpose_any_kernel", "transpose_words_kernel",
                 "transpose_u16_kernel", "gather_words_kernel", "group4_kernel", "pad_channels_kernel", "split3_kernel",
                 "fold_filter_kernel", "selftest_fastq_kernel", "selftest_div_kernel", "selftest_packed_kernel"):
        if cand in name:
            fn = cand
            break
    fam[fn][0] += 1
    fam[fn][2] += len(re.findall(r"/\*[0-9a-f]{4,}\*/\s+\S", f))
    for k in keys:
        n = len(re.findall(re.escape(k), f))
        fam[fn][1][k] += n
        tot[k] += n
print("SASS summary of megengine_b200/lib/libb200cb.so (cuobjdump -sass, sm_100a only), %d kernels" % len(funcs))
print("totals:", " ".join("%s=%d" % (k, tot[k]) for k in keys))
print("\nper kernel family (instances, SASS instructions, mnemonic counts summed over the instances):")
for fn, (cnt, c, n) in sorted(fam.items(), key=lambda x: -x[1][2]):
    print("  %-26s x%-3d %7d instr | %s" % (fn, cnt, n, " ".join("%s=%d" % (k, c[k]) for k in keys if c[k])))
