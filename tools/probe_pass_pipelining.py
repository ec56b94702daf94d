"""Upper bound of pass pipelining: two integrators (two films, two workspaces, two streams) rendering alternate passes of
the same scene concurrently vs one integrator rendering all of them.  If the pair is not faster, pipelining passes across
the pass boundary cannot pay."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import zillumgl_b200 as zl
import bench as B

res = {}
for wl in sys.argv[1:] or ["rungholt"]:
    scene, w, h, kind, desc, _ = B.build_scene(zl, wl, 0, 0)
    K = 16
    def timed(integs):
        for it in integs:
            for _ in range(3): it.renderOnePass()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(K):
            integs[k % len(integs)].renderOnePass()
        torch.cuda.synchronize()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K
    one = timed([B.make_integrator(zl, scene, kind, w, h, None, 1)])
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    a = zl.__dict__[{"path": "NaivePathIntegrator", "light": "LightPathIntegrator", "triple": "TriplePathIntegrator"}[kind]]
    def mk(stream):
        it = a(scene, w, h, stream=stream.cuda_stream)
        it.mParam.kernelVariant = 1
        if kind == "light": it.mParam.threadBlocksOnePass = (w * h + 1535) // 1536
        if kind == "triple": it.mParam.LPTBlocksOnePass = 64
        return it
    two = timed([mk(s1), mk(s2)])
    s3 = torch.cuda.Stream()
    three = timed([mk(s1), mk(s2), mk(s3)])
    res[wl] = {"one_stream_ms_per_pass": one, "two_streams_ms_per_pass": two, "three_streams_ms_per_pass": three}
    print(wl, res[wl], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "probe_pass_pipelining.json"), "w"), indent=1)
