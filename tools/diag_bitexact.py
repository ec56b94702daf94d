"""Diagnostic: how many lanes / pixels of the CUDA path differ IN BITS from the CPU oracle, per
KAT op and per integrator film (NaN == NaN).  Run on the GPU box: python tools/diag_bitexact.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import zillumgl_b200 as zl  # noqa: E402
from conftest import get_scene  # noqa: E402


def bits(a):
    return np.asarray(a).astype(np.int32).view(np.float32)


def mism(g, r):
    bad = (g.view(np.uint32) != r.view(np.uint32)) & ~(np.isnan(g) & np.isnan(r))
    # +0 / -0 are reported separately
    zero = bad & (g == 0) & (r == 0)
    return int((bad & ~zero).any(axis=-1).sum()), int(zero.any(axis=-1).sum())


def params(s, w, h, **kw):
    p = zl.ZlRenderParams()
    p.camera = s.camera(); p.camera.asp = w / h
    p.filmW, p.filmH = w, h
    p.maxDepth, p.sampleLight, p.lightPortion, p.sampler = 4, 1, 0.5, 1
    p.spp, p.freeCounter = 3, 4
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def unit(rng, k):
    v = rng.normal(size=(k, 3)).astype(np.float32)
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def kat():
    rng = np.random.default_rng(8)
    n = 1 << 14
    for scene, w, h, mats in (("cornell", 64, 48, [0, 1, 3, 4]), ("default", 64, 36, [1, 2]), ("sponza_light", 64, 36, [0, 5, 8, 9])):
        s, o = get_scene(scene, w, h)
        if not s.device:
            s.upload()
        p = params(s, w, h)
        p.camera.lensRadius, p.camera.focalDist = 0.05, 3.0
        inp = rng.random((n, 6), dtype=np.float32)
        g, r = zl.debug_eval(s, p, zl.KAT["CAMERA_RAY"], inp, 6), o.debug_eval(p, zl.KAT["CAMERA_RAY"], inp, 6)
        print(scene, "CAMERA_RAY", mism(g, r))
        ref = (rng.random((n, 3), dtype=np.float32) * 2 - 1) + np.array([0, 0, 1], np.float32)
        inp = np.concatenate([ref, rng.random((n, 2), dtype=np.float32)], axis=1)
        g, r = zl.debug_eval(s, p, zl.KAT["CAMERA_II"], inp, 10), o.debug_eval(p, zl.KAT["CAMERA_II"], inp, 10)
        print(scene, "CAMERA_II", mism(g, r))
        for mat in mats:
            nrm, wo, wi = unit(rng, n), unit(rng, n), unit(rng, n)
            for mode in (0, 1):
                ev = np.zeros((n, 14), np.float32)
                ev[:, 0] = bits([mat])[0]; ev[:, 1] = bits([-1])[0]
                ev[:, 4:7], ev[:, 7:10], ev[:, 10:13], ev[:, 13] = wo, wi, nrm, bits([mode])[0]
                g, r = zl.debug_eval(s, p, zl.KAT["BSDF_EVAL"], ev, 4), o.debug_eval(p, zl.KAT["BSDF_EVAL"], ev, 4)
                sm = np.zeros((n, 15), np.float32)
                sm[:, 0] = bits([mat])[0]; sm[:, 1] = bits([-1])[0]
                sm[:, 4:7], sm[:, 7:10], sm[:, 10] = wo, nrm, bits([mode])[0]
                sm[:, 11:14] = rng.random((n, 3), dtype=np.float32)
                sm[:, 14] = bits(rng.integers(0, 2 ** 31, n))
                g2, r2 = zl.debug_eval(s, p, zl.KAT["BSDF_SAMPLE"], sm, 9), o.debug_eval(p, zl.KAT["BSDF_SAMPLE"], sm, 9)
                print(scene, "mat", mat, "mode", mode, "EVAL", mism(g, r), "SAMPLE", mism(g2, r2))
                bad = (g2.view(np.uint32) != r2.view(np.uint32)) & ~(np.isnan(g2) & np.isnan(r2))
                if bad.any():
                    i = np.nonzero(bad.any(axis=1))[0][0]
                    print("   first:", sm[i], g2[i], r2[i])
    s, o = get_scene("rungholt_small", 64, 36)
    if not s.device:
        s.upload()
    p = params(s, 64, 36, envRotation=0.7)
    d = unit(rng, n)
    g, r = zl.debug_eval(s, p, zl.KAT["ENV_LE"], d, 4), o.debug_eval(p, zl.KAT["ENV_LE"], d, 4)
    print("ENV_LE", mism(g, r))
    u = rng.random((n, 4), dtype=np.float32)
    g, r = zl.debug_eval(s, p, zl.KAT["ENV_SAMPLE"], u, 4), o.debug_eval(p, zl.KAT["ENV_SAMPLE"], u, 4)
    print("ENV_SAMPLE", mism(g, r))
    s, o = get_scene("sponza_light", 64, 36)
    p = params(s, 64, 36)
    nl = s.info["nLightTriangles"]
    lid = rng.integers(0, nl, n)
    u = rng.random((n, 4), dtype=np.float32)
    inp = np.concatenate([bits(lid).reshape(-1, 1), u], axis=1)
    g, r = zl.debug_eval(s, p, zl.KAT["LIGHT_SAMPLE_LE"], inp, 11), o.debug_eval(p, zl.KAT["LIGHT_SAMPLE_LE"], inp, 11)
    print("LIGHT_SAMPLE_LE", mism(g, r))
    x = (rng.random((n, 3), dtype=np.float32) - 0.5) * np.array([30, 10, 8], np.float32) + np.array([0, 0, 4.5], np.float32)
    inp = np.concatenate([x, rng.random((n, 5), dtype=np.float32)], axis=1)
    g, r = zl.debug_eval(s, p, zl.KAT["SAMPLE_LIGHT_ENV"], inp, 7), o.debug_eval(p, zl.KAT["SAMPLE_LIGHT_ENV"], inp, 7)
    print("SAMPLE_LIGHT_ENV", mism(g, r))


def films():
    for kind, cls in (("path", zl.NaivePathIntegrator), ("triple", zl.TriplePathIntegrator), ("light", zl.LightPathIntegrator)):
        for name, w, h in (("cornell", 64, 48), ("default", 64, 36), ("rungholt_small", 64, 36), ("sponza_light", 48, 27)):
            s, o = get_scene(name, w, h)
            if not s.device:
                s.upload()
            for variant in ((0, 1, 2) if kind == "path" else (0, 1)):
                integ = cls(s, w, h)
                integ.mParam.kernelVariant = variant
                if kind == "light":
                    integ.mParam.threadBlocksOnePass = 2
                if kind == "triple":
                    integ.mParam.LPTBlocksOnePass = 1
                ref = np.zeros((h, w, 4), np.float32)
                for _ in range(4):
                    if kind == "path":
                        o.path_pass(integ.params(), ref)
                    elif kind == "light":
                        o.light_pass(integ.params(), ref)
                    else:
                        o.triple_pt_pass(integ.params(0), ref)
                        o.triple_lpt_pass(integ.params(1), ref)
                    integ.renderOnePass()
                img = integ.getFrame(1.0)[..., :3]
                a, b = np.ascontiguousarray(img), np.ascontiguousarray(ref[..., :3])
                nb, nz = mism(a, b)
                rel = np.abs(a - b).max() / (np.abs(b).max() + 1e-30)
                print(f"film {kind:6s} {name:15s} v{variant}: pixels differing {nb} (+{nz} zero-sign) of {w * h}, max abs diff / max = {rel:.3e}")


if __name__ == "__main__":
    kat()
    films()
