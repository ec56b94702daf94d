"""GPU image parity per integrator (through the host Integrator classes and the C ABI) with
identical seeds / sample streams on both sides.  The MIS path tracer's film (one owner per
pixel, plain adds) is BIT-IDENTICAL to the oracle's; the light and triple tracers splat with
float atomics, so their films agree up to summation order only (tolerance stated in
_assert_splat_film).  The north star's gate, relMSE < 1e-3 at 4096 spp per integrator
(SURVEY.md §8d gate 2), is asserted on top of that."""
import numpy as np
import pytest

from conftest import get_scene, rel_mse

pytestmark = pytest.mark.gpu


def _scene(name, w, h):
    s, o = get_scene(name, w, h)
    if not s.device:
        s.upload()
    return s, o


def _render_pair(zl, kind, name, w, h, passes, **params):
    s, o = _scene(name, w, h)
    cls = {"path": zl.NaivePathIntegrator, "light": zl.LightPathIntegrator, "triple": zl.TriplePathIntegrator}[kind]
    integ = cls(s, w, h)
    for k, v in params.items():
        setattr(integ.mParam, k, v)
    ref = np.zeros((h, w, 4), np.float32)
    for _ in range(passes):
        if kind == "path":
            o.path_pass(integ.params(), ref)
        elif kind == "light":
            o.light_pass(integ.params(), ref)
        else:
            o.triple_pt_pass(integ.params(0), ref)
            o.triple_lpt_pass(integ.params(1), ref)
        integ.renderOnePass()
    scale = integ.trueScale()
    return integ.getFrame()[..., :3], ref[..., :3] * scale, integ


def _assert_same_film(img, ref):
    """bit for bit (NaN never reaches the film: hasNan() filters, path_integ_naive.glsl:172)"""
    a, b = np.ascontiguousarray(img, np.float32), np.ascontiguousarray(ref, np.float32)
    bad = a.view(np.uint32) != b.view(np.uint32)
    assert not bad.any(), (int(bad.any(axis=-1).sum()), a[bad][:4], b[bad][:4])


def _assert_splat_film(img, ref, terms=1):
    """Films that receive imageAtomicAdd splats (light_path_integ.glsl:36-42): equal up to the order of the
    float additions.  Bound used: |a - b| <= 8 eps sqrt(terms) (|b| + max|b| / 256), eps = 2^-24, `terms` the
    order of magnitude of the splats a pixel receives — the error of a random-order sum grows with sqrt(n)."""
    a, b = np.asarray(img, np.float64), np.asarray(ref, np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(fin, np.isfinite(a))
    tol = 8 * 2.0 ** -24 * np.sqrt(max(terms, 1)) * (np.abs(b[fin]) + np.abs(b[fin]).max() / 256)
    bad = np.abs(a[fin] - b[fin]) > tol
    assert not bad.any(), (int(bad.sum()), np.abs(a[fin] - b[fin]).max(), np.abs(b[fin]).max())


@pytest.mark.parametrize("name,w,h", [("cornell", 64, 48), ("default", 64, 36), ("rungholt_small", 64, 36), ("sponza_light", 48, 27)])
def test_path_tracer_single_pass_matches_per_pixel(name, w, h, zl):
    """Same seeds, same Sobol dimensions, same arithmetic: the film is the oracle's bit for bit."""
    img, ref, _ = _render_pair(zl, "path", name, w, h, passes=2)
    _assert_same_film(img, ref)
    assert ref.mean() > 1e-4


@pytest.mark.parametrize("name,w,h,spp", [("cornell", 32, 24, 4096), ("default", 48, 27, 512), ("rungholt_small", 48, 27, 256), ("sponza_light", 32, 18, 256)])
def test_path_tracer_converged_relmse(name, w, h, spp, zl):
    img, ref, integ = _render_pair(zl, "path", name, w, h, passes=spp)
    assert integ.curSample == spp
    _assert_same_film(img, ref)
    assert rel_mse(img, ref) < 1e-3          # the north star's gate (trivially: the films are identical)
    assert not np.isnan(img).any()


@pytest.mark.parametrize("kw", [dict(russianRoulette=1), dict(sampleLight=0), dict(maxDepth=1), dict(maxDepth=8, russianRoulette=1),
                                dict(lightEnvUniformSample=1, lightPortion=0.3)])
def test_path_tracer_parameter_variants(kw, zl):
    img, ref, _ = _render_pair(zl, "path", "rungholt_small", 48, 27, passes=96, **kw)
    _assert_same_film(img, ref)


def test_path_tracer_hash_sampler(zl):
    s, _ = _scene("cornell", 48, 36)
    s.set_sampler(0)
    try:
        img, ref, _ = _render_pair(zl, "path", "cornell", 48, 36, passes=256)
    finally:
        s.set_sampler(1)
    _assert_same_film(img, ref)


@pytest.mark.parametrize("name,w,h,passes,blocks", [("cornell", 32, 24, 4096, 1), ("default", 48, 27, 512, 2), ("sponza_light", 32, 18, 256, 1)])
def test_light_tracer_converged_relmse(name, w, h, passes, blocks, zl):
    # 4096 passes x 1536 paths over 768 pixels = 8192 light paths per pixel
    img, ref, integ = _render_pair(zl, "light", name, w, h, passes=passes, threadBlocksOnePass=blocks)
    _assert_splat_film(img, ref, terms=passes * blocks * 1536 * 4 / (w * h))
    assert rel_mse(img, ref) < 1e-3
    assert ref.sum() > 0 and not np.isnan(img).any()


def test_light_tracer_russian_roulette_and_depth(zl):
    img, ref, _ = _render_pair(zl, "light", "cornell", 48, 36, passes=512, threadBlocksOnePass=2, russianRoulette=1, maxDepth=6)
    _assert_splat_film(img, ref, terms=512 * 2 * 1536 * 6 / (48 * 36))
    assert rel_mse(img, ref) < 1e-3


@pytest.mark.parametrize("name,w,h,passes", [("cornell", 32, 24, 4096), ("default", 48, 27, 384), ("sponza_light", 32, 18, 192)])
def test_triple_tracer_converged_relmse(name, w, h, passes, zl):
    img, ref, _ = _render_pair(zl, "triple", name, w, h, passes=passes, LPTBlocksOnePass=1)
    _assert_splat_film(img, ref, terms=passes * (1 + 1536 * 4 / (w * h)))
    assert rel_mse(img, ref) < 1e-3
    assert not np.isnan(img).any()


def test_triple_tracer_loops_and_blocks(zl):
    img, ref, _ = _render_pair(zl, "triple", "cornell", 48, 36, passes=256, LPTBlocksOnePass=2, LPTLoopsPerPass=2, russianRoulette=1)
    _assert_splat_film(img, ref, terms=256 * (1 + 4 * 1536 * 4 / (48 * 36)))
    assert rel_mse(img, ref) < 1e-3


def test_integrator_bookkeeping_matches_reference_semantics(zl):
    """resultScale() keeps the reference's off-by-one (Integrator.h:79, LightPath.cpp:113,133);
    trueScale() is 1 / true sample count."""
    s, _ = _scene("cornell", 32, 24)
    pt = zl.NaivePathIntegrator(s, 32, 24)
    assert pt.mParam.maxDepth == 4 and pt.mParam.sampleLight == 1 and pt.mParam.russianRoulette == 0 and pt.mParam.maxSample == 64
    for _ in range(3):
        pt.renderOnePass()
    assert pt.curSample == 3 and pt.resultScale() == pytest.approx(1 / 4) and pt.trueScale() == pytest.approx(1 / 3)
    pt.mParam.finiteSample, pt.mParam.maxSample = 1, 4
    for _ in range(10):
        pt.renderOnePass()
    assert pt.curSample == 5                                  # stops after maxSample + 1 passes (App. B #21)
    lt = zl.LightPathIntegrator(s, 32, 24)
    assert lt.mParam.threadBlocksOnePass == 32
    lt.renderOnePass()
    per = 32 * 1536 / (32 * 24)
    assert lt.mParam.samplePerPixel == pytest.approx(2 * per) and lt.trueScale() == pytest.approx(1 / per)
    tp = zl.TriplePathIntegrator(s, 32, 24)
    assert tp.mParam.LPTBlocksOnePass == 64 and tp.mParam.LPTLoopsPerPass == 1
    assert tp.params(1).scale == pytest.approx(32 * 24 / (64 * 1536))
    pt.reset()
    assert pt.curSample == 0 and float(np.abs(pt.getFrame(1.0)[..., :3]).max()) == 0.0


def test_sample_index_sharding_equals_single_device(zl):
    """Multi-GPU partition exercised on one device: N shards rendering passes
    g, g+N, ... and summed equal the single-integrator film up to FP32 summation order."""
    s, _ = _scene("default", 64, 36)
    spp, world = 16, 4
    whole = zl.NaivePathIntegrator(s, 64, 36)
    for _ in range(spp):
        whole.renderOnePass()
    full = whole.getFrame(1.0)
    acc = np.zeros_like(full)
    for g in range(world):
        part = zl.NaivePathIntegrator(s, 64, 36)
        part.setSampleShard(g, world)
        for _ in range(spp // world):
            part.renderOnePass()
        acc += part.getFrame(1.0)
    assert np.allclose(acc[..., :3], full[..., :3], rtol=1e-5, atol=1e-6)


def test_external_film_memory(zl):
    torch = pytest.importorskip("torch")
    s, _ = _scene("cornell", 32, 24)
    film = torch.zeros((24, 32, 4), dtype=torch.float32, device="cuda")
    a = zl.NaivePathIntegrator(s, 32, 24, external_film_ptr=film.data_ptr())
    b = zl.NaivePathIntegrator(s, 32, 24)
    for _ in range(4):
        a.renderOnePass(); b.renderOnePass()
    zl.synchronize()
    assert np.array_equal(film.cpu().numpy()[..., :3], b.getFrame(1.0)[..., :3])


def test_instrumented_pass_counts_match_oracle(zl):
    """The separately compiled counting build visits exactly the nodes / triangles the oracle
    visits for the same pass (the roofline byte model rests on these counters)."""
    s, o = _scene("default", 64, 36)
    integ = zl.NaivePathIntegrator(s, 64, 36)
    integ.mParam.sampleLight = 0          # without NEE every ray of the pass is decided by exact arithmetic + the sampler
    p = integ.params()
    ref = np.zeros((36, 64, 4), np.float32)
    st = o.path_pass(p, ref)
    c = zl.counted_pass(s, integ.film, p, 0)
    assert c["paths"] == 64 * 36 == st["paths"]
    assert abs(c["rays"] - st["rays"]) <= 0.002 * st["rays"]
    assert abs(c["nodes"] - st["nodeVisits"]) <= 0.005 * st["nodeVisits"]
    assert abs(c["tris"] - st["triTests"]) <= 0.005 * st["triTests"]
    img = integ.getFrame(1.0)
    _assert_same_film(img[..., :3], ref[..., :3])


@pytest.mark.parametrize("name,w,h,kw", [
    ("cornell", 64, 48, {}), ("default", 61, 35, {}), ("rungholt_small", 64, 36, {}), ("sponza_light", 50, 27, {}),
    ("rungholt_small", 48, 27, dict(russianRoulette=1)), ("default", 48, 27, dict(sampleLight=0)),
    ("cornell", 48, 36, dict(maxDepth=1)), ("rungholt_small", 48, 27, dict(maxDepth=8, russianRoulette=1)),
    ("sponza_light", 48, 27, dict(lightEnvUniformSample=1, lightPortion=0.3))])
def test_wavefront_variant_is_bit_identical_to_megakernel(name, w, h, kw, zl):
    """Variant 1 (wavefront: primary / shade / trace / resolve stages over compacted queues, NEE
    shadow rays deferred) performs the same arithmetic per path in the same order as the
    megakernel, so the accumulated film must be equal BIT FOR BIT, including odd film sizes."""
    import os
    s, _ = _scene(name, w, h)
    frames = []
    # megakernel; wavefront default; without ray sorting; with the regenerating (ballot/popc refill) trace kernel;
    # with the round-based refill kernel (from bounce 0, 5-step rounds so that rays span several rounds); finer sort cells
    for variant, sort, simple, extra in ((0, "1", "3", {}), (1, "1", "3", {}), (1, "0", "3", {}), (1, "1", "0", {}),
                                         (1, "1", "3", {"ZL_WF_TRACE_LOOP": "4", "ZL_WF_REFILL_FROM": "0", "ZL_WF_ROUND_STEPS": "5", "ZL_WF_REFILL_AT": "3"}),
                                         (1, "1", "3", {"ZL_WF_SORT_BITS": "6"}),
                                         (1, "1", "3", {"ZL_OCTANT_WALK": "0"}),          # general packed walk only (no octant-specialised loops)
                                         (1, "1", "3", {"ZL_WF_TRACE_LOOP": "5"}),        # two rays per lane (wfTraceDualKernel), sorted queues
                                         (1, "0", "3", {"ZL_WF_TRACE_LOOP": "5"}),        # ... unsorted: mixed octants take traverseDual<-1>
                                         (1, "1", "3", {"ZL_WF_TRACE_LOOP": "0", "ZL_WF_TRACE_PIPE": "1"}),      # chunk heads software-pipelined (cp.async queue entries, L2 prefetch)
                                         (1, "1", "3", {"ZL_WF_TRACE_LOOP": "0", "ZL_NODE_POLICY": "1", "ZL_STATE_POLICY": "1"}),   # general (non-lean) instantiation, eviction priorities
                                         (1, "1", "3", {"ZL_WF_TRACE_LOOP": "0", "ZL_WF_TRACE_CTAS_PER_SM": "3"})):     # a trace grid smaller than the SMs can hold
        os.environ["ZL_WF_SORT"], os.environ["ZL_WF_TRACE_SIMPLE"] = sort, simple
        for k in ("ZL_WF_TRACE_LOOP", "ZL_WF_REFILL_FROM", "ZL_WF_ROUND_STEPS", "ZL_WF_REFILL_AT", "ZL_WF_SORT_BITS", "ZL_OCTANT_WALK", "ZL_WF_TRACE_PIPE",
                  "ZL_NODE_POLICY", "ZL_STATE_POLICY", "ZL_WF_TRACE_CTAS_PER_SM"):
            os.environ.pop(k, None)
        os.environ.update(extra)
        integ = zl.NaivePathIntegrator(s, w, h)
        integ.mParam.kernelVariant = variant
        for k, v in kw.items():
            setattr(integ.mParam, k, v)
        for _ in range(6):
            integ.renderOnePass()
        frames.append(integ.getFrame(1.0))
    for k in ("ZL_WF_SORT", "ZL_WF_TRACE_SIMPLE", "ZL_WF_TRACE_LOOP", "ZL_WF_REFILL_FROM", "ZL_WF_ROUND_STEPS", "ZL_WF_REFILL_AT", "ZL_WF_SORT_BITS", "ZL_OCTANT_WALK",
              "ZL_WF_TRACE_PIPE", "ZL_NODE_POLICY", "ZL_STATE_POLICY", "ZL_WF_TRACE_CTAS_PER_SM"):
        os.environ.pop(k, None)
    assert frames[0][..., :3].max() > 0
    for f in frames[1:]:
        assert np.array_equal(frames[0].view(np.uint32), f.view(np.uint32))


@pytest.mark.parametrize("bits", ["0", "1", "3", "4", "5", "6", "7", "8", "31", "-2"])
def test_sort_bits_over_and_beyond_the_allowed_range(bits, zl):
    """ZL_WF_SORT_BITS (bits per axis of the Morton cell in the ray-sort key) is clamped to [4, 7]: below 4 a histogram would not be a
    whole number of 8192-bin scan tiles, above 7 the key would not fit.  Every setting, in range or not, must give the megakernel's
    film bit for bit (the sort only reorders the queue) with the ray sort forced on."""
    import os
    w, h = 64, 36
    s, _ = _scene("rungholt_small", w, h)
    ref = zl.NaivePathIntegrator(s, w, h)
    ref.mParam.kernelVariant = 0
    os.environ["ZL_WF_SORT"], os.environ["ZL_WF_SORT_BITS"] = "1", bits
    try:
        integ = zl.NaivePathIntegrator(s, w, h)          # a new film: the workspace (and its histograms) is sized with this setting
        integ.mParam.kernelVariant = 1
        for _ in range(5):
            ref.renderOnePass(); integ.renderOnePass()
        a, b = ref.getFrame(1.0), integ.getFrame(1.0)
    finally:
        os.environ.pop("ZL_WF_SORT", None); os.environ.pop("ZL_WF_SORT_BITS", None)
    assert a[..., :3].max() > 0 and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_wavefront_variant_hash_sampler_and_relmse(zl):
    img, ref, _ = _render_pair(zl, "path", "cornell", 48, 36, passes=64, kernelVariant=1)
    _assert_same_film(img, ref)


def test_headless_cli_writes_the_same_image(zl, tmp_path):
    """zillum_render (headless driver, EXR/PFM output) = the Integrator classes driven from C++."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(zl.__file__), "host", "zillum_render")
    out = tmp_path / "cornell.pfm"
    r = subprocess.run([exe, "builtin:cornell", "--integrator", "path", "--spp", "8", "--size", "64x48", "--out", str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    with open(out, "rb") as f:
        assert f.readline().strip() == b"PF"
        w, h = map(int, f.readline().split())
        scale = float(f.readline())
        data = np.frombuffer(f.read(), "<f4" if scale < 0 else ">f4").reshape(h, w, 3)
    s, _ = _scene("cornell", 64, 48)
    integ = zl.NaivePathIntegrator(s, 64, 48)
    for _ in range(8):
        integ.renderOnePass()
    assert (w, h) == (64, 48)
    assert np.allclose(data, integ.getFrame()[..., :3], rtol=1e-6, atol=1e-7)
    exr = tmp_path / "cornell.exr"
    r = subprocess.run([exe, "builtin:cornell", "--integrator", "light", "--spp", "2", "--size", "32x24", "--out", str(exr)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and exr.stat().st_size > 32 * 24 * 12


def test_pipelined_frame_readback_equals_blocking_one(zl):
    """getFrameAsync()/waitFrame(): resolve on the render stream, D2H on a copy stream overlapping the next
    passes; every frame must equal what the blocking getFrame() returns at the same point."""
    torch = pytest.importorskip("torch")
    s, _ = _scene("default", 64, 36)
    a, b = zl.NaivePathIntegrator(s, 64, 36), zl.NaivePathIntegrator(s, 64, 36)
    bufs = [torch.empty((36, 64, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    want = []
    for k in range(5):
        b.renderOnePass(); want.append(b.getFrame(1.0).copy())
    got = []
    for k in range(5):
        a.renderOnePass()
        if k > 0:
            a.waitFrame(); got.append(bufs[(k - 1) % 2].numpy().copy())
        a.getFrameAsync(bufs[k % 2].data_ptr(), 1.0)
    a.waitFrame(); got.append(bufs[4 % 2].numpy().copy())
    for g, w_ in zip(got, want):
        assert np.array_equal(g.view(np.uint32), w_.view(np.uint32))


@pytest.mark.parametrize("name,w,h,kw", [
    ("cornell", 48, 36, dict(threadBlocksOnePass=2)), ("default", 48, 27, dict(threadBlocksOnePass=1)),
    ("sponza_light", 32, 18, dict(threadBlocksOnePass=1)), ("cornell", 32, 24, dict(threadBlocksOnePass=3, russianRoulette=1, maxDepth=6)),
    ("rungholt_small", 48, 27, dict(threadBlocksOnePass=2, maxDepth=1)), ("cornell", 32, 24, dict(threadBlocksOnePass=1, maxDepth=0))])
def test_light_tracer_wavefront_equals_megakernel(name, w, h, kw, zl):
    """Variant 1 of the light pass (generate / shade per material type / sort / trace-and-splat stages) traces
    the same light paths with the same random numbers as the megakernel; only the order in which the float
    atomics land differs, so the films agree to FP32 summation order."""
    s, _ = _scene(name, w, h)
    frames = []
    for variant in (0, 1):
        integ = zl.LightPathIntegrator(s, w, h)
        integ.mParam.kernelVariant = variant
        for k, v in kw.items():
            setattr(integ.mParam, k, v)
        for _ in range(24):
            integ.renderOnePass()
        frames.append(integ.getFrame(1.0)[..., :3].astype(np.float64))
    assert frames[0].max() > 0
    scale = np.abs(frames[0]).max()
    assert np.abs(frames[0] - frames[1]).max() <= 2e-5 * scale + 1e-6, np.abs(frames[0] - frames[1]).max() / scale


@pytest.mark.parametrize("name,w,h,kw", [
    ("cornell", 48, 36, {}), ("default", 61, 35, {}), ("sponza_light", 50, 27, {}), ("rungholt_small", 48, 27, dict(russianRoulette=1)),
    ("cornell", 32, 24, dict(maxDepth=1)), ("rungholt_small", 48, 27, dict(maxDepth=7, russianRoulette=1))])
def test_triple_camera_pass_wavefront_is_bit_identical(name, w, h, kw, zl):
    """PT pass of the triple tracer alone (LPTBlocksOnePass = 0 disables the light pass): one owner per pixel,
    same arithmetic in the same order => variant 1 equals the megakernel bit for bit."""
    s, _ = _scene(name, w, h)
    frames = []
    for variant in (0, 1):
        integ = zl.TriplePathIntegrator(s, w, h)
        integ.mParam.kernelVariant, integ.mParam.LPTBlocksOnePass = variant, 0
        for k, v in kw.items():
            setattr(integ.mParam, k, v)
        for _ in range(6):
            integ.renderOnePass()
        frames.append(integ.getFrame(1.0))
    assert frames[0][..., :3].max() > 0
    assert np.array_equal(frames[0].view(np.uint32), frames[1].view(np.uint32))


@pytest.mark.parametrize("name,w,h,kw", [
    ("cornell", 48, 36, dict(LPTBlocksOnePass=1)), ("default", 48, 27, dict(LPTBlocksOnePass=1, LPTLoopsPerPass=3)),
    ("sponza_light", 32, 18, dict(LPTBlocksOnePass=2, russianRoulette=1)), ("rungholt_small", 48, 27, dict(LPTBlocksOnePass=1, LPTLoopsPerPass=2, maxDepth=2))])
def test_triple_tracer_wavefront_equals_megakernel(name, w, h, kw, zl):
    """Both passes: the light pass splats with float atomics, so equality holds to FP32 summation order."""
    s, _ = _scene(name, w, h)
    frames = []
    for variant in (0, 1):
        integ = zl.TriplePathIntegrator(s, w, h)
        integ.mParam.kernelVariant = variant
        for k, v in kw.items():
            setattr(integ.mParam, k, v)
        for _ in range(12):
            integ.renderOnePass()
        frames.append(integ.getFrame(1.0)[..., :3].astype(np.float64))
    assert frames[0].max() > 0
    scale = np.abs(frames[0]).max()
    assert np.abs(frames[0] - frames[1]).max() <= 2e-5 * scale + 1e-6, np.abs(frames[0] - frames[1]).max() / scale


def test_stage_timing_reports_the_trace_kernel(zl):
    """zl_stage_timing_*: CUDA-event time and launch count per stage of the wavefront pass; the
    traversal stage is launched once per bounce 0..maxDepth and nothing is recorded when disabled."""
    s, _ = _scene("rungholt_small", 64, 36)
    integ = zl.NaivePathIntegrator(s, 64, 36)
    integ.mParam.kernelVariant = 1
    integ.renderOnePass()
    zl.stage_timing_enable(True)
    try:
        integ.renderOnePass()
        integ.renderOnePass()
        st = zl.stage_timing_read()
    finally:
        zl.stage_timing_enable(False)
    depth = int(integ.mParam.maxDepth)
    assert st["trace"][1] == 2 * (depth + 1) and st["trace"][0] > 0.0
    assert st["generate"][1] == 2 and st["resolve"][1] == 2 * (depth + 1)
    assert st["shade"][1] >= 2 * depth and st["megakernel"][1] == 0
    integ.mParam.kernelVariant = 0
    zl.stage_timing_enable(True)
    try:
        integ.renderOnePass()
        st = zl.stage_timing_read()
    finally:
        zl.stage_timing_enable(False)
    assert st["megakernel"][1] == 1 and st["trace"][1] == 0
    integ.renderOnePass()
    assert sum(v[1] for v in zl.stage_timing_read().values()) == 0


def test_full_size_properties_rungholt_c5(zl):
    """BASELINE config C5 at full size (6,291,456 triangles, 3840x2160, the bench workload): properties that do not need
    the CPU to trace 8.3 M paths — a strided subset of the 4K primary rays against the oracle bit for bit, any-hit
    consistent with closest-hit, the wavefront pass bit-identical to the megakernel pass, film accumulation linear in
    the passes, sample-index shards summing to the unsharded film, and the device-threaded MTBVH the host-threaded one."""
    from conftest import rel_mse
    import oracle_lib
    w, h = 3840, 2160
    s = zl.Scene.builtin("rungholt", w, h)
    s.set_device_bvh(True)                          # the bench path: BVH::build and the MTBVH threading both on the device
    s.flatten()
    s.upload()
    assert s.device_prep_times()["bvh_levels"] > 10
    assert s.info["numTriangles"] == 6291456
    o = oracle_lib.OracleScene(s.desc)              # threads its own hit table (the scene carries none)
    p = zl.ZlRenderParams()
    p.camera = s.camera(); p.camera.asp = w / h
    p.filmW, p.filmH = w, h
    rs = zl.RaySet.primary(p)
    rs.trace(s)
    ids, t = rs.download()
    rays = rs.rays()
    sub = np.arange(0, ids.size, 499)
    rid, rt = o.trace_rays(rays[sub])
    assert np.array_equal(ids[sub], rid) and np.array_equal(t[sub], rt)
    hit = np.nonzero(ids >= 0)[0][::17]
    assert hit.size > 100000
    lo, hi = (t[hit] * 0.999).astype(np.float32), (t[hit] * 1.001).astype(np.float32)
    assert zl.trace_rays(s, rays[hit], anyhit=True, tmax=lo)[0].sum() <= 0.001 * hit.size
    assert zl.trace_rays(s, rays[hit], anyhit=True, tmax=hi)[0].mean() > 0.999
    # device-built, device-threaded node records = reference texels of the oracle's own tree and table, on windows of every face
    n = s.info["bvhSize"]
    ob, ot = oracle_lib.build_bvh(s.array("vertices"), s.array("indices"))
    ot = ot.reshape(6, n, 3); ob = ob.reshape(n, 6)
    for f in range(6):
        for first in (0, n // 2, n - 4096):
            b, l = s.read_nodes(f, first, 4096)
            assert np.array_equal(l, ot[f, first:first + 4096, 1:3])
            ref = ob[ot[f, first:first + 4096, 0]]
            assert np.all((b == ref) | ((b == 0) & (ref == 0)))                      # zeros may differ in sign (zl_bvh_build.cuh)
    # one pass each way; two passes; two shards
    def render(variant, passes, shard=None):
        integ = zl.NaivePathIntegrator(s, w, h)
        integ.mParam.kernelVariant = variant
        if shard:
            integ.setSampleShard(*shard)
        for _ in range(passes):
            integ.renderOnePass()
        return integ.getFrame(1.0)
    mega, wave = render(0, 1), render(1, 1)
    assert np.array_equal(mega.view(np.uint32), wave.view(np.uint32)) and wave[..., :3].max() > 0
    two = render(1, 2)
    second = render(1, 1, shard=(1, 2))             # pass index 1 alone
    assert np.array_equal(two[..., :3], wave[..., :3] + second[..., :3])       # r0 + r1 in the same order as the film's accumulation
    # a sample of film rows of pass 0 against the oracle (identical sample streams): relMSE gate of SURVEY §8d
    integ = zl.NaivePathIntegrator(s, w, h)
    ref = np.zeros((h, w, 4), np.float32)
    o.path_pass(integ.params(), ref, 7, h, 240)     # 9 rows spread over the film
    rows = np.arange(7, h, 240)
    _assert_same_film(wave[rows][..., :3], ref[rows][..., :3])                 # bit for bit, like the small scenes


@pytest.mark.parametrize("name,w,h,kw", [
    ("cornell", 64, 48, {}), ("rungholt_small", 61, 35, {}), ("sponza_light", 50, 27, dict(russianRoulette=1)),
    ("default", 48, 27, dict(maxDepth=8, russianRoulette=1)), ("rungholt_small", 48, 27, dict(maxDepth=1))])
def test_pipelined_passes_are_bit_identical(name, w, h, kw, zl):
    """kernelVariant 2 (two passes in flight on internal streams, every film write and read ordered on one film stream):
    the film must equal the megakernel's and the sequential wavefront's bit for bit at every point it is observed —
    after an odd and an even number of passes, through getFrame, getFrameAsync (snapshot between passes), postProcess,
    across reset(), and when variants are mixed on one film."""
    import torch
    s, _ = _scene(name, w, h)
    def make(variant):
        integ = zl.NaivePathIntegrator(s, w, h)
        integ.mParam.kernelVariant = variant
        for k, v in kw.items():
            setattr(integ.mParam, k, v)
        return integ
    ref, pipe = make(0), make(2)
    pinned = [torch.empty((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    for n in range(1, 8):
        ref.renderOnePass(); pipe.renderOnePass()
        if n in (1, 2, 5):
            assert np.array_equal(ref.getFrame(1.0).view(np.uint32), pipe.getFrame(1.0).view(np.uint32)), f"after {n} passes"
        if n in (3, 6):      # snapshot while the next pass is already being launched
            pipe.getFrameAsync(pinned[n % 2].data_ptr(), 1.0)
            expect = ref.getFrame(1.0)
            ref.renderOnePass(); pipe.renderOnePass()
            pipe.waitFrame()
            assert np.array_equal(pinned[n % 2].numpy().view(np.uint32), expect.view(np.uint32)), f"snapshot after {n} passes"
    a, _ = ref.postProcess("filmic")
    b, _ = pipe.postProcess("filmic")
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    # reset, then mixed variants on the same film: pipelined, sequential wavefront, megakernel, pipelined
    ref.reset(); pipe.reset()
    for variant in (2, 2, 1, 0, 2, 2, 2):
        pipe.mParam.kernelVariant = variant
        ref.renderOnePass(); pipe.renderOnePass()
    pipe.flush()
    assert np.array_equal(ref.getFrame(1.0).view(np.uint32), pipe.getFrame(1.0).view(np.uint32))
    assert ref.getFrame(1.0)[..., :3].max() > 0


def test_pipelined_passes_external_film_and_flush(zl):
    """A torch-owned film (the multi-GPU path): after flush() the tensor holds every pass, as the sequential schedule leaves it."""
    import torch
    w, h = 64, 36
    s, _ = _scene("rungholt_small", w, h)
    films = [torch.zeros((h, w, 4), dtype=torch.float32, device="cuda") for _ in range(2)]
    out = []
    for variant, film in zip((1, 2), films):
        integ = zl.NaivePathIntegrator(s, w, h, external_film_ptr=film.data_ptr())
        integ.mParam.kernelVariant = variant
        for _ in range(5):
            integ.renderOnePass()
        integ.flush()
        torch.cuda.synchronize()
        out.append(film.cpu().numpy().copy())
        del integ
    assert np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32)) and out[0][..., :3].max() > 0


@pytest.mark.parametrize("variant", [1, 2])
def test_two_readbacks_in_flight(variant, zl):
    """getFrameAsync may be called twice before waitFrame (own staging buffers, FIFO completion); a third call takes over the
    oldest slot.  Every frame must be the snapshot after exactly the passes launched before it."""
    import torch
    w, h = 64, 48
    s, _ = _scene("cornell", w, h)
    ref = zl.NaivePathIntegrator(s, w, h); ref.mParam.kernelVariant = 0
    integ = zl.NaivePathIntegrator(s, w, h); integ.mParam.kernelVariant = variant
    pinned = [torch.empty((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(4)]
    expect = []
    for k in range(4):
        ref.renderOnePass(); expect.append(ref.getFrame(1.0))
        integ.renderOnePass()
        if k == 3:
            integ.waitFrame()                       # frame 0 (oldest); frames 1 and 2 stay in flight
            assert np.array_equal(pinned[0].numpy().view(np.uint32), expect[0].view(np.uint32))
        integ.getFrameAsync(pinned[k].data_ptr(), 1.0)       # k = 2: third read-back while two are in flight -> takes over slot of frame 0
        if k == 2:
            pass
    for _ in range(3):
        integ.waitFrame()
    integ.waitFrame()                               # none in flight: no-op
    for k in range(4):
        assert np.array_equal(pinned[k].numpy().view(np.uint32), expect[k].view(np.uint32)), f"frame {k}"


def test_rgb_readback_equals_rgba_frame(zl):
    """getFrameAsync(channels=3): the packed RGB frame is the RGBA frame without its constant alpha."""
    import torch
    w, h = 61, 35
    s, _ = _scene("default", w, h)
    integ = zl.NaivePathIntegrator(s, w, h); integ.mParam.kernelVariant = 2
    rgb = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()
    for _ in range(3):
        integ.renderOnePass()
    integ.getFrameAsync(rgb.data_ptr(), 0.5, channels=3)
    integ.waitFrame()
    rgba = integ.getFrame(0.5)
    assert np.array_equal(rgb.numpy().view(np.uint32), np.ascontiguousarray(rgba[..., :3]).view(np.uint32)) and np.all(rgba[..., 3] == 1.0)


def test_light_tracer_pipelined_passes(zl):
    """Light tracer, kernelVariant 2: two passes splat concurrently (float atomics: equal up to summation order), and a frame
    read between passes contains exactly the passes launched before it."""
    import torch
    w, h = 64, 48
    s, _ = _scene("cornell", w, h)
    def make(variant):
        integ = zl.LightPathIntegrator(s, w, h)
        integ.mParam.kernelVariant = variant
        integ.mParam.threadBlocksOnePass = 4
        return integ
    seq, pipe = make(1), make(2)
    pinned = [torch.empty((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    for n in range(1, 7):
        seq.renderOnePass(); pipe.renderOnePass()
        if n in (2, 5):      # snapshot, then the next pass is launched while the read is in flight
            pipe.getFrameAsync(pinned[n % 2].data_ptr(), 1.0)
            expect = seq.getFrame(1.0)
            seq.renderOnePass(); pipe.renderOnePass()
            pipe.waitFrame()
            got = pinned[n % 2].numpy()
            assert rel_mse(got, expect) < 1e-10 and abs(float(got[..., :3].sum()) / float(expect[..., :3].sum()) - 1.0) < 1e-5, f"snapshot after {n} passes"
    a, b = seq.getFrame(1.0), pipe.getFrame(1.0)
    assert a[..., :3].max() > 0 and rel_mse(b, a) < 1e-10
    pipe.reset(); seq.reset()
    for _ in range(3):
        seq.renderOnePass(); pipe.renderOnePass()
    pipe.flush()
    assert rel_mse(pipe.getFrame(1.0), seq.getFrame(1.0)) < 1e-10


def test_triple_tracer_pipelined_passes(zl):
    """Triple tracer, kernelVariant 2: camera pass + light pass per chain, two pass pairs in flight.  Equal to the sequential
    wavefront schedule up to the summation order of the splats; frame reads contain whole passes only; converges to the oracle."""
    import torch
    w, h = 48, 27
    s, o = _scene("sponza_light", w, h)
    def make(variant):
        integ = zl.TriplePathIntegrator(s, w, h)
        integ.mParam.kernelVariant = variant
        integ.mParam.LPTBlocksOnePass = 2
        integ.mParam.LPTLoopsPerPass = 2
        return integ
    seq, pipe = make(1), make(2)
    pinned = [torch.empty((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    for n in range(1, 8):
        seq.renderOnePass(); pipe.renderOnePass()
        if n in (2, 5):
            pipe.getFrameAsync(pinned[n % 2].data_ptr(), 1.0)
            expect = seq.getFrame(1.0)
            seq.renderOnePass(); pipe.renderOnePass()
            pipe.waitFrame()
            assert rel_mse(pinned[n % 2].numpy(), expect) < 1e-10, f"snapshot after {n} passes"
    a, b = seq.getFrame(1.0), pipe.getFrame(1.0)
    assert a[..., :3].max() > 0 and rel_mse(b, a) < 1e-10
    # against the oracle, same streams: the usual single-pass bar
    ref = np.zeros((h, w, 4), np.float32)
    chk = make(2)
    for _ in range(4):
        o.triple_pt_pass(chk.params(0), ref)
        o.triple_lpt_pass(chk.params(1), ref)
        chk.renderOnePass()
    assert rel_mse(chk.getFrame()[..., :3], ref[..., :3] * chk.trueScale()) < 5e-3


@pytest.mark.parametrize("name,w,h,kw", [
    ("cornell", 64, 48, {}), ("rungholt_small", 61, 35, {}), ("sponza_light", 50, 27, dict(russianRoulette=1)),
    ("default", 48, 27, dict(maxDepth=8, russianRoulette=1))])
def test_graph_replayed_passes_are_bit_identical(name, w, h, kw, zl):
    """kernelVariant 3 (pass 1 plain launches, pass 2 stream-captured, pass 3.. one cudaGraphLaunch each with uSpp / uFreeCounter
    read from the device pair the graph's first node writes): the film equals the megakernel's bit for bit after every pass,
    across reset() (the counters restart), a parameter change (new graph) and mixed with the other variants on one film."""
    s, _ = _scene(name, w, h)
    def make(variant):
        integ = zl.NaivePathIntegrator(s, w, h)
        integ.mParam.kernelVariant = variant
        for k, v in kw.items():
            setattr(integ.mParam, k, v)
        return integ
    ref, gr = make(0), make(3)
    before = zl.launch_count()
    for n in range(1, 9):
        ref.renderOnePass(); gr.renderOnePass()
        assert np.array_equal(ref.getFrame(1.0).view(np.uint32), gr.getFrame(1.0).view(np.uint32)), f"after {n} passes"
    assert zl.launch_count() > before
    ref.reset(); gr.reset()
    for variant in (3, 3, 1, 2, 3, 0, 3, 3):
        gr.mParam.kernelVariant = variant
        ref.renderOnePass(); gr.renderOnePass()
    gr.flush()
    assert np.array_equal(ref.getFrame(1.0).view(np.uint32), gr.getFrame(1.0).view(np.uint32))
    ref.mParam.maxDepth = gr.mParam.maxDepth = 2          # baked into the graph: must be re-captured
    ref.reset(); gr.reset()
    for _ in range(5):
        ref.renderOnePass(); gr.renderOnePass()
    assert np.array_equal(ref.getFrame(1.0).view(np.uint32), gr.getFrame(1.0).view(np.uint32))
    assert ref.getFrame(1.0)[..., :3].max() > 0


def test_graph_replayed_light_and_triple_passes(zl):
    """kernelVariant 3 for the splatting integrators: same kernels in the same order as variant 1 => equal up to the summation
    order of the float atomics (and the triple tracer's camera pass bit for bit when the light pass is off)."""
    w, h = 64, 48
    s, _ = _scene("cornell", w, h)
    films = []
    for variant in (1, 3):
        integ = zl.LightPathIntegrator(s, w, h)
        integ.mParam.kernelVariant, integ.mParam.threadBlocksOnePass = variant, 4
        for _ in range(7):
            integ.renderOnePass()
        films.append(integ.getFrame(1.0))
    assert films[0][..., :3].max() > 0 and rel_mse(films[1], films[0]) < 1e-10
    s, _ = _scene("sponza_light", 48, 27)
    for blocks, exact in ((2, False), (0, True)):
        films = []
        for variant in (1, 3):
            integ = zl.TriplePathIntegrator(s, 48, 27)
            integ.mParam.kernelVariant, integ.mParam.LPTBlocksOnePass, integ.mParam.LPTLoopsPerPass = variant, blocks, 2
            for _ in range(6):
                integ.renderOnePass()
            films.append(integ.getFrame(1.0))
        assert films[0][..., :3].max() > 0
        if exact:
            assert np.array_equal(films[0].view(np.uint32), films[1].view(np.uint32))
        else:
            assert rel_mse(films[1], films[0]) < 1e-10


@pytest.mark.parametrize("kind", ["path", "light", "triple"])
def test_converged_4096_passes_on_the_sponza_class_scene(kind, zl):
    """The north star's image gate at its stated sample count on the 262 k-triangle scene (BASELINE configs C3 / C4), 96x54 film:
    4096 passes per integrator, CUDA against the oracle with identical sample streams — relMSE < 1e-3, and in fact the path tracer's
    film is identical and the splat films agree to summation order."""
    w, h = 96, 54
    kw = {"light": dict(threadBlocksOnePass=1), "triple": dict(LPTBlocksOnePass=1)}.get(kind, {})
    img, ref, integ = _render_pair(zl, kind, "sponza_light", w, h, passes=4096, **kw)
    assert rel_mse(img, ref) < 1e-3
    if kind == "path":
        assert integ.curSample == 4096
        _assert_same_film(img, ref)
    else:
        _assert_splat_film(img, ref, terms=4096 * (1 + 1536 * 4 / (w * h)))
    assert ref.mean() > 1e-3 and not np.isnan(img).any()


def test_cornell_light_tracer_c2_film_size_nonfinite_pixels_match(zl):
    """BASELINE config C2 at its film size (1920x1080, default 32 blocks per pass): a few passes; every non-finite film value the
    reference's arithmetic produces (light_path_integ.glsl:118 filters NaN results before a splat, not infinite ones) sits in the
    same pixel on both sides, and the finite part agrees to summation order."""
    w, h = 1920, 1080
    img, ref, _ = _render_pair(zl, "light", "cornell", w, h, passes=6)
    assert np.array_equal(np.isfinite(img), np.isfinite(ref))
    assert np.array_equal(np.isnan(img), np.isnan(ref))
    fin = np.isfinite(ref)
    assert fin.mean() > 0.999 and ref[fin].sum() > 0
    a, b = np.where(fin, img, 0.0), np.where(fin, ref, 0.0)
    _assert_splat_film(a, b, terms=6 * 32 * 1536 * 4 / (w * h) + 4)
