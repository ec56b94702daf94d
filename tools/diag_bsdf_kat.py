"""Diagnostic (GPU box): where do BSDF_SAMPLE lanes of CUDA and the oracle disagree?"""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import zillumgl_b200 as zl
from conftest import get_scene
import test_gpu_kat as T

out = {}
for scene, mats in (("cornell", [0, 1, 3, 4]), ("default", [1, 2])):
    w, h = (64, 48) if scene == "cornell" else (64, 36)
    s, o, p = T._setup(zl, scene, w, h)
    rng = np.random.default_rng(8)
    n = 1 << 16
    for mat in mats:
        def unit(k):
            v = rng.normal(size=(k, 3)).astype(np.float32)
            return v / np.linalg.norm(v, axis=1, keepdims=True)
        nrm, wo = unit(n), unit(n)
        mtype = int(np.asarray(s.array("materials")).reshape(-1, 16)[mat, 13:14].view(np.int32)[0])
        for mode in (0,):
            sm = np.zeros((n, 15), np.float32)
            sm[:, 0] = T._bits([mat])[0]; sm[:, 1] = T._bits([-1])[0]
            sm[:, 4:7], sm[:, 7:10], sm[:, 10] = wo, nrm, T._bits([mode])[0]
            sm[:, 11:14] = rng.random((n, 3), dtype=np.float32)
            sm[:, 14] = T._bits(rng.integers(0, 2 ** 31, n))
            g, r = T._both(zl, s, o, p, "BSDF_SAMPLE", sm, 9)
            same_flag = g[:, 8].view(np.uint32) == r[:, 8].view(np.uint32)
            close = np.isclose(g[:, :8], r[:, :8], rtol=2e-4, atol=2e-6).all(axis=1) | (np.isnan(g[:, :8]) & np.isnan(r[:, :8])).any(axis=1)
            ok = same_flag & close
            cosn = (wo * nrm).sum(axis=1)
            up = cosn > 0
            d = {"type": mtype, "agree_all": float(ok.mean()), "agree_up": float(ok[up].mean()), "agree_down": float(ok[~up].mean()),
                 "agree_up_cos>0.05": float(ok[cosn > 0.05].mean()),
                 "bad_up_cos_quantiles": np.quantile(cosn[up & ~ok], [0, .5, .9, 1]).tolist() if (up & ~ok).any() else None,
                 "bad_up_pdf_quantiles": np.quantile(r[up & ~ok, 3], [0, .5, .9, 1]).tolist() if (up & ~ok).any() else None,
                 "bad_up_examples": [dict(inp=sm[i, 4:14].tolist(), g=g[i, :8].tolist(), r=r[i, :8].tolist()) for i in np.nonzero(up & ~ok)[0][:4]]}
            out[f"{scene}/mat{mat}"] = d
            print(scene, mat, {k: v for k, v in d.items() if k != "bad_up_examples"}, flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "diag_bsdf.json"), "w"), indent=1)
