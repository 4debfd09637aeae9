"""run one gpu_diag case by name: python tools/one_case.py <stage> <case>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.gpu_diag import stage_cases
from tests import _cases as K, _gpu as G, _oracle as O
from megengine_b200 import _capi as C
stage, name = sys.argv[1], sys.argv[2]
for n, d, opt in stage_cases(stage):
    if n != name:
        continue
    src, flt, bias, z = K.gen_inputs(d, 1234)
    want = O.oracle_convbias(d, src, flt, bias, z)
    want_f32 = K.fp32_reference(d, src, flt, bias, z) if d.dst_dtype in (C.DT_BF16, C.DT_F16) else None
    import ctypes, numpy as np, torch
    for variant in (0, 1):
        rc, plan = G.plan_for(d, opt.get("algo"))
        plan.reserved[6] = variant
        tsrc, tflt, tbias, tz = G.to_dev(src), G.to_dev(flt), G.to_dev(bias), G.to_dev(z)
        tdst = torch.full((O.dst_elems(d),), 0x55, dtype=torch.int16, device="cuda")
        tws = torch.empty(int(plan.workspace_bytes), dtype=torch.uint8, device="cuda") if plan.workspace_bytes else None
        G.launch(plan, tsrc, tflt, tbias, tz, tdst, tws)
        torch.cuda.synchronize()
        got = tdst.cpu().numpy().view(np.uint16)
        ok, info = G.compare(d, got, want, want_f32)
        print(n, "variant", variant, "mode", plan.a_mode, ok, {k: v for k, v in info.items() if k != "first_bad"})
        if not ok:
            g = O.to_f32(got, d.dst_dtype).reshape(d.n, d.oh, d.ow, d.oc)
            w = want_f32.reshape(d.n, d.oh, d.ow, d.oc)
            bad = ~(O.megdnn_rel_err(g, w) <= 1e-2)
            print(" bad per image", bad.reshape(d.n, -1).sum(1), "bad per oh (img0)", bad[0].reshape(d.oh, -1).sum(1))
            print(" bad per ow (img0)", bad[0].transpose(1, 0, 2).reshape(d.ow, -1).sum(1))
            print(" bad per channel (first 16)", bad.reshape(-1, d.oc).sum(0)[:16])
            print(" sample got/want img0 oh0:", g[0, 0, :4, :4].tolist(), w[0, 0, :4, :4].tolist())
