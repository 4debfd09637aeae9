"""Staged GPU diagnostics (run under gpurun): each stage runs in its own subprocess with a timeout
so that a hung kernel cannot take the later stages with it.  Writes gpurun_out/diag.jsonl.

    python -m tests.gpu_diag [stage ...]
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def stage_cases(stage):
    from megengine_b200 import _capi as C
    from . import _cases as K
    md = K.make_desc
    if stage == "gemm_tiled":       # pointwise conv == plain GEMM through tiled TMA
        return [("pw_bf16_1tile", md("bf16", 1, 64, 8, 16, 64, 1, bias_mode=C.BIAS_NONE), {}),
                ("pw_bf16_bias", md("bf16", 1, 64, 8, 16, 64, 1), {}),
                ("pw_bf16_k256_oc128", md("bf16", 2, 256, 16, 16, 128, 1, nonline=1), {}),
                ("pw_bf16_oc256", md("bf16", 2, 128, 16, 16, 256, 1, nonline=1), {}),
                ("pw_bf16_oc512_x256", md("bf16", 2, 128, 16, 16, 512, 1, nonline=1), {"algo": 0}),
                ("pw_bf16_oc512_x128", md("bf16", 2, 128, 16, 16, 512, 1, nonline=1), {"algo": 1}),
                ("pw_bf16_oc512_x64", md("bf16", 2, 128, 16, 16, 512, 1, nonline=1), {"algo": 2}),
                ("pw_bf16_mtail", md("bf16", 3, 64, 7, 7, 64, 1, nonline=1), {}),
                ("pw_bf16_z", md("bf16", 2, 64, 14, 14, 256, 1, nonline=1, has_z=True), {}),
                ("pw_f16_z", md("f16", 2, 64, 14, 14, 256, 1, nonline=1, has_z=True), {}),
                ("pw_bf16_ic16", md("bf16", 2, 16, 14, 14, 64, 1, nonline=1), {}),
                ("pw_bf16_ic144_oc24", md("bf16", 2, 144, 14, 14, 24, 1, nonline=0), {})]
    if stage == "gemm_im2col":      # same GEMMs forced through the im2col descriptor
        return [("pwi_bf16_1tile", md("bf16", 1, 64, 8, 16, 64, 1, bias_mode=C.BIAS_NONE), {"im2col": True}),
                ("pwi_bf16_mtail", md("bf16", 3, 64, 7, 7, 64, 1, nonline=1), {"im2col": True}),
                ("pwi_bf16_k256", md("bf16", 2, 256, 16, 16, 128, 1, nonline=1), {"im2col": True}),
                ("pwi_bf16_s2", md("bf16", 2, 64, 16, 16, 64, 1, stride=2), {})]
    if stage == "conv":
        return [("c3_bf16_p1", md("bf16", 2, 64, 14, 14, 64, 3, pad=1, nonline=1), {}),
                ("c3_bf16_p0", md("bf16", 2, 64, 14, 14, 64, 3, pad=0, nonline=1), {}),
                ("c3_bf16_s2", md("bf16", 2, 64, 15, 15, 64, 3, stride=2, pad=1, nonline=1), {}),
                ("c3_bf16_asym", md("bf16", 2, 64, 12, 17, 64, 3, pad=1, pad_w=0, stride=1, stride_w=2), {}),
                ("c5_bf16_p2", md("bf16", 2, 32, 13, 13, 64, 5, pad=2, nonline=3), {}),
                ("c7_bf16_s2_ic8", md("bf16", 1, 8, 32, 32, 64, 7, stride=2, pad=3, nonline=1), {}),
                ("c3_bf16_z_oc256", md("bf16", 2, 128, 14, 14, 256, 3, pad=1, nonline=1, has_z=True), {}),
                ("c3_bf16_dil2", md("bf16", 2, 64, 14, 14, 64, 3, pad=2, dilate=2), {}),
                ("c3_bf16_flip", md("bf16", 2, 64, 9, 9, 64, 3, pad=1, mode=1), {}),
                ("c3_f16", md("f16", 2, 64, 14, 14, 128, 3, pad=1, nonline=1, has_z=True), {}),
                ("c3_bf16_big", md("bf16", 8, 128, 28, 28, 128, 3, pad=1, nonline=1), {}),
                # narrow K-blocks (SWIZZLE_64B / 32B) and the W-folded stride-2 stem
                ("c3_bf16_ic32_sw64", md("bf16", 2, 32, 14, 14, 64, 3, pad=1, nonline=1), {}),
                ("c3_bf16_ic16_sw32", md("bf16", 2, 16, 14, 14, 64, 3, pad=1, nonline=1, has_z=True), {}),
                ("c3_bf16_ic24_sw64_tail", md("bf16", 2, 24, 14, 14, 64, 3, pad=1), {}),
                ("c3_bf16_ic8_s1", md("bf16", 2, 8, 14, 14, 64, 3, pad=1, nonline=1), {}),
                ("fold_c7_bf16_ic8", md("bf16", 2, 8, 30, 30, 64, 7, stride=2, pad=3, nonline=1), {}),
                ("fold_c3_bf16_ic8_p1", md("bf16", 2, 8, 18, 18, 64, 3, stride=2, pad=1, nonline=1), {}),
                ("fold_c3_bf16_ic8_p0", md("bf16", 2, 8, 18, 18, 64, 3, stride=2, pad=0), {}),
                ("fold_c5_f16_ic8_p2", md("f16", 2, 8, 20, 20, 128, 5, stride=2, pad=2, nonline=1), {}),
                ("fold_c2_bf16_ic8_p0", md("bf16", 1, 8, 16, 16, 64, 2, stride=2, pad=0), {}),
                ("stem_c3_bf16_wide_ow150", md("bf16", 1, 8, 20, 300, 64, 3, stride=2, pad=1, nonline=1), {}),
                ("stem_c3_bf16_oc32_z", md("bf16", 2, 8, 18, 18, 32, 3, stride=2, pad=1, nonline=1, has_z=True), {}),
                ("stem_c7_f16_ic8_hswish", md("f16", 1, 8, 34, 34, 64, 7, stride=2, pad=3, nonline=3), {}),
                # stem4: 8-byte pixels (bf16 / fp16 IC = 4, int8 IC = 8), a pixel pair per 16-byte unit
                ("stem4_c7_bf16_ic4", md("bf16", 2, 4, 30, 30, 64, 7, stride=2, pad=3, nonline=1), {}),
                ("stem4_c7_bf16_ic4_wide_z", md("bf16", 1, 4, 12, 300, 64, 7, stride=2, pad=3, nonline=1, has_z=True), {}),
                ("stem4_c3_f16_ic4_p1_oc32", md("f16", 2, 4, 18, 18, 32, 3, stride=2, pad=1, nonline=1), {}),
                ("stem4_c5_bf16_ic4_p2", md("bf16", 2, 4, 20, 20, 64, 5, stride=2, pad=2), {}),
                ("stem4_c3_bf16_ic4_p0", md("bf16", 3, 4, 18, 22, 64, 3, stride=2, pad=0, nonline=3), {}),
                ("fold4_c7_bf16_ic4_oc128", md("bf16", 2, 4, 30, 30, 128, 7, stride=2, pad=3, nonline=1), {}),
                # rows mode (resident input rows + resident filter)
                ("rows_c3_bf16_56", md("bf16", 2, 64, 56, 56, 64, 3, pad=1, nonline=1), {}),
                ("rows_c3_bf16_7x7_n3", md("bf16", 3, 64, 7, 7, 64, 3, pad=1, nonline=1), {}),
                ("rows_c5_f16_oc32", md("f16", 2, 64, 20, 13, 32, 5, pad=2, nonline=3), {}),
                ("rows_c3_bf16_padh0", md("bf16", 2, 64, 12, 17, 64, 3, pad=0, pad_w=1), {}),
                ("rows_c1x3_bf16", md("bf16", 2, 64, 9, 30, 64, 1, fw=3, pad=0, pad_w=1, nonline=1), {}),
                ("rows_c3_bf16_ic128_oc16", md("bf16", 1, 128, 10, 10, 16, 3, pad=1), {}),
                ("rows64_c3_bf16_ic32", md("bf16", 2, 32, 28, 28, 64, 3, pad=1, nonline=1), {}),
                ("rows64_c3_f16_ic32_oc48", md("f16", 3, 32, 9, 11, 48, 3, pad=1, nonline=3), {}),
                ("nofold_c3_bf16_ic8_oddw", md("bf16", 2, 8, 17, 17, 64, 3, stride=2, pad=1), {})]
    if stage == "int8":
        sc = (1.2, 1.3, None, 0.7, 19.99)
        return [("q_pw", md("qs8", 2, 128, 14, 14, 64, 1, nonline=0, scales=sc), {}),
                ("q_pw_oc256_relu", md("qs8", 2, 128, 14, 14, 256, 1, nonline=1, scales=sc), {}),
                ("q_c3_z_relu", md("qs8", 2, 64, 14, 14, 128, 3, pad=1, nonline=1, has_z=True, scales=sc), {}),
                ("q_c3_z_hswish", md("qs8", 2, 64, 14, 14, 128, 3, pad=1, nonline=3, has_z=True,
                                     scales=(0.11, 0.23, None, 0.37, 0.41)), {"int_range": 127}),
                ("q_c3_s2_nobias", md("qs8", 3, 32, 15, 15, 64, 3, stride=2, pad=1, nonline=1,
                                      bias_mode=C.BIAS_NONE, scales=sc), {}),
                ("q_pw_qs32", md("qs8", 2, 64, 14, 14, 64, 1, dst_kind="qs32"), {}),
                ("q_c3_stress", md("qs8", 2, 256, 14, 14, 256, 3, pad=1, nonline=1, has_z=True,
                                   scales=(0.013, 0.021, None, 0.05, 0.9)), {"int_range": 127, "bias_range": 5000}),
                ("q_stem4_c7_ic8", md("qs8", 2, 8, 30, 30, 64, 7, stride=2, pad=3, nonline=1, scales=sc),
                 {"int_range": 127}),
                ("q_stem4_c3_ic8_z", md("qs8", 2, 8, 18, 18, 64, 3, stride=2, pad=1, nonline=1, has_z=True, scales=sc), {}),
                ("q_fold_c7_ic16", md("qs8", 2, 16, 30, 30, 64, 7, stride=2, pad=3, nonline=1, scales=sc),
                 {"int_range": 127}),
                ("q_c3_ic32_sw32", md("qs8", 2, 32, 14, 14, 64, 3, pad=1, nonline=1, has_z=True, scales=sc), {}),
                ("q_c3_ic64_sw64_oc512", md("qs8", 2, 64, 14, 14, 512, 3, pad=1, nonline=3, has_z=True,
                                            scales=(0.11, 0.23, None, 0.37, 0.41)), {"int_range": 127}),
                ("q_pw_oc192_tail", md("qs8", 3, 128, 7, 7, 192, 1, nonline=1, has_z=True, scales=sc), {"algo": 1}),
                ("q_denormal_dst_scale", md("qs8", 1, 64, 7, 7, 64, 1, scales=(0.1, 0.2, None, 0.3, 1e-39)), {}),
                # rows mode with 64-byte channel rows (SWIZZLE_64B slabs): QuantizedS8 IC = 64 (ResNet-50 l1 3x3)
                ("q_rows64_c3_56", md("qs8", 2, 64, 56, 56, 64, 3, pad=1, nonline=1, scales=sc), {"int_range": 127}),
                ("q_rows64_c3_7x7_n3_hswish", md("qs8", 3, 64, 7, 7, 64, 3, pad=1, nonline=3,
                                                 scales=(0.11, 0.23, None, 0.37, 0.41)), {"int_range": 127}),
                ("q_rows64_c5_oc32", md("qs8", 2, 64, 20, 13, 32, 5, pad=2, nonline=0, scales=sc), {}),
                ("q_rows64_c1x3_padw", md("qs8", 2, 64, 9, 30, 64, 1, fw=3, pad=0, pad_w=1, nonline=1, scales=sc), {}),
                ("q_rows128_c3_ic128", md("qs8", 2, 128, 12, 17, 64, 3, pad=1, nonline=1, scales=sc), {"int_range": 127})]
    if stage == "dw":
        out = []
        for kind in ("bf16", "f16", "f32", "qs8"):
            out += [("dw3s1_" + kind, K.dw_desc(kind, 2, 32, 14, 14, 1, has_z=True), {}),
                    ("dw3s2_" + kind, K.dw_desc(kind, 2, 96, 15, 15, 2), {}),
                    ("dw5s1_" + kind, K.dw_desc(kind, 2, 48, 9, 9, 1, f=5, nonline=3), {}),
                    ("dw3s1_c144_" + kind, K.dw_desc(kind, 3, 144, 7, 7, 1), {})]
        out.append(("dw7_generic_bf16", K.dw_desc("bf16", 2, 32, 12, 12, 1, f=7), {}))
        # TMA bandwidth kernel: channel tails (64-blocks + 16 / 32 / 40-channel tail launches), wide and
        # partial tiles, z, stride 2, odd sizes; channel counts the TMA path cannot take (C % 8 != 0)
        out += [("dw3_tma_c144_56", K.dw_desc("bf16", 2, 144, 56, 56, 1, has_z=True), {}),
                ("dw3_tma_c96_s2_57", K.dw_desc("bf16", 2, 96, 57, 57, 2), {}),
                ("dw3_tma_c232_s2", K.dw_desc("f16", 2, 232, 14, 14, 2, nonline=3), {}),
                ("dw3_tma_c24_s2", K.dw_desc("bf16", 3, 24, 30, 30, 2), {}),
                ("dw3_tma_c16", K.dw_desc("bf16", 2, 16, 20, 20, 1, has_z=True), {}),
                ("dw3_tma_c8_wide", K.dw_desc("f16", 1, 8, 9, 130, 1), {}),
                # small images: several whole images per tile (ragged last tile: N not a multiple)
                ("dw3_tma_7x7_n3", K.dw_desc("bf16", 3, 64, 7, 7, 1), {}),
                ("dw3_tma_7x7_n11_z", K.dw_desc("f16", 11, 136, 7, 7, 1, has_z=True), {}),
                ("dw3_tma_14s2_n5", K.dw_desc("bf16", 5, 72, 14, 14, 2, nonline=3), {}),
                ("dw3_tma_5x6_n9_c32", K.dw_desc("bf16", 9, 32, 5, 6, 1), {}),
                # pixels that are not 16-byte multiples: bulk-copied row blocks when the rows are 16-byte
                # multiples (several tiles per image, ragged spans, stride 2, z), else the register front end
                ("dw3_bulk_c58", K.dw_desc("bf16", 2, 58, 28, 28, 1), {}),
                ("dw3_bulk_c116_s2_z", K.dw_desc("f16", 2, 116, 28, 28, 2, has_z=True), {}),
                ("dw3_bulk_c116_56", K.dw_desc("bf16", 2, 116, 56, 56, 1, nonline=3), {}),
                ("dw3_bulk_c36_22x30_s2_z", K.dw_desc("bf16", 3, 36, 22, 30, 2, has_z=True), {}),
                ("dw3_bulk_c58_56_s2", K.dw_desc("f16", 5, 58, 56, 56, 2), {}),
                ("dw3_bulk_c6_3x8", K.dw_desc("bf16", 2, 6, 3, 8, 1), {}),
                # no padding (valid convolution) through the three front ends
                ("dw3_tma_pad0", K.make_desc("bf16", 2, 64, 12, 12, 64, 3, stride=1, pad=0, group=64, nonline=1), {}),
                ("dw3_tma_pad0_s2", K.make_desc("f16", 3, 40, 13, 13, 40, 3, stride=2, pad=0, group=40, nonline=1), {}),
                ("dw3_bulk_pad0_s2", K.make_desc("bf16", 2, 36, 11, 16, 36, 3, stride=2, pad=0, group=36, nonline=1), {}),
                ("dw3_bulk_pad0", K.make_desc("bf16", 2, 12, 9, 8, 12, 3, stride=1, pad=0, group=12, nonline=1, has_z=True), {}),
                ("dw3_reg_pad0", K.make_desc("bf16", 2, 58, 9, 9, 58, 3, stride=1, pad=0, group=58, nonline=1), {}),
                ("dw3_reg_c58_14", K.dw_desc("bf16", 2, 58, 14, 14, 1), {}),
                ("dw3_reg_c116_15_s2_z", K.dw_desc("f16", 2, 116, 15, 15, 2, has_z=True), {})]
        return out
    raise SystemExit("unknown stage " + stage)


def analyse(d, got, want_f32_or_int, info):
    """where are the wrong outputs? rows (pixels) vs columns (channels) histogram."""
    from megengine_b200 import _capi as C
    from . import _oracle as O
    oc = d.oc
    if d.dst_dtype in (C.DT_QS8, C.DT_QS32):
        bad = (got != want_f32_or_int)
    else:
        g = O.to_f32(got, d.dst_dtype)
        bad = ~(O.megdnn_rel_err(g, want_f32_or_int) <= info["tol"])
    bad = bad.reshape(-1, oc)
    rows = np.nonzero(bad.any(axis=1))[0]
    cols = np.nonzero(bad.any(axis=0))[0]
    info["bad_rows"] = int(rows.size)
    info["bad_cols"] = int(cols.size)
    info["rows_head"] = rows[:24].tolist()
    info["cols_head"] = cols[:24].tolist()
    info["shape"] = [int(bad.shape[0]), int(oc)]


def run_stage(stage):
    import torch
    from megengine_b200 import _capi as C
    from . import _cases as K, _gpu as G, _oracle as O
    assert torch.cuda.is_available()
    for name, d, opt in stage_cases(stage):
        rec = {"stage": stage, "case": name, "desc": K.desc_str(d)}
        try:
            src, flt, bias, z = K.gen_inputs(d, 1234, int_range=opt.get("int_range", 8),
                                             bias_range=opt.get("bias_range", 50))
            want = O.oracle_convbias(d, src, flt, bias, z)
            want_f32 = None
            if d.dst_dtype in (C.DT_BF16, C.DT_F16):
                want_f32 = K.fp32_reference(d, src, flt, bias, z)
            t0 = time.time()
            got, plan = G.run_gpu(d, src, flt, bias, z, algo=opt.get("algo"), force_im2col=opt.get("im2col", False))
            rec["algo"] = C.load().b200cb_algo_name(plan.algo).decode()
            rec["stages"] = plan.stages
            rec["a_mode"] = 1 if opt.get("im2col") else plan.a_mode
            ok, info = G.compare(d, got, want, want_f32)
            rec["ok"] = bool(ok)
            rec.update(info)
            if not ok:
                analyse(d, got, want_f32 if want_f32 is not None else
                        (O.to_f32(want, d.dst_dtype) if d.dst_dtype == C.DT_F32 else want), rec)
            rec["sec"] = round(time.time() - t0, 3)
        except Exception as e:  # noqa: BLE001
            rec["ok"] = False
            rec["error"] = repr(e)[:400]
        print(json.dumps(rec), flush=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    stages = sys.argv[1:] or ["gemm_tiled", "gemm_im2col", "conv", "int8", "dw"]
    if os.environ.get("B200CB_DIAG_CHILD"):
        run_stage(stages[0])
        return
    path = os.path.join(OUT, "diag.jsonl")
    with open(path, "a") as f:
        for st in stages:
            env = dict(os.environ, B200CB_DIAG_CHILD="1")
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, "-m", "tests.gpu_diag", st], env=env, cwd=ROOT,
                                   capture_output=True, text=True, timeout=240)
                out, err, code = r.stdout, r.stderr, r.returncode
            except subprocess.TimeoutExpired as e:
                out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                err, code = "TIMEOUT", -9
            f.write(out)
            summary = {"stage": st, "exit": code, "sec": round(time.time() - t0, 1), "stderr_tail": err[-600:]}
            f.write(json.dumps(summary) + "\n")
            f.flush()
            lines = [json.loads(l) for l in out.splitlines() if l.startswith("{")]
            nok = sum(1 for l in lines if l.get("ok"))
            print("[%s] %d/%d ok, exit %d, %.1fs" % (st, nok, len(lines), code, time.time() - t0))
            for l in lines:
                if not l.get("ok"):
                    print("   FAIL", json.dumps(l)[:900])
            if code != 0:
                print("   stderr:", err[-600:])


if __name__ == "__main__":
    main()
