"""Display stage: post_proc.glsl tone mapping + gamma and the PNG screenshot writer.
CPU: the oracle's restatement against closed-form values, the PNG writer against a zlib decode.
GPU: zl_film_postprocess against the oracle on rendered films."""
import struct
import zlib

import numpy as np
import pytest

import oracle_lib as O


def _read_png(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(raw):
        n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
        body = raw[pos + 8:pos + 8 + n]
        (crc,) = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])
        assert crc == zlib.crc32(typ + body) & 0xffffffff, "chunk CRC"
        chunks.append((typ, body))
        pos += 12 + n
    assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
    w, h, depth, ctype, comp, flt, lace = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, ctype, comp, flt, lace) == (8, 2, 0, 0, 0)
    data = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(h, 1 + 3 * w)
    assert np.all(data[:, 0] == 0)
    return data[:, 1:].reshape(h, w, 3)


def test_png_writer_round_trip(zl, tmp_path):
    rng = np.random.default_rng(3)
    for w, h in ((1, 1), (7, 5), (257, 131), (640, 360)):           # the last one spans several 64 KiB stored blocks
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        p = tmp_path / f"t{w}x{h}.png"
        assert zl.write_png(p, img)
        # rows are given bottom-up (film order) and written top row first (stbi_flip_vertically_on_write, Application.cpp:376)
        assert np.array_equal(_read_png(p), img[::-1])


def test_oracle_post_proc_closed_form():
    film = np.zeros((1, 6, 4), np.float32)
    film[0, :, :3] = np.array([0.0, 0.18, 1.0, 11.2 / 1.6, -3.0, 1e35], np.float32)[:, None]
    g = 1.0 / 2.2
    # no tone mapping: pow(clamp(x * scale, 0, 1e30), 1/2.2)
    out, out8 = O.post_proc(film, 0.5, 0)
    x = np.clip(film[0, :, 0].astype(np.float64) * 0.5, 0.0, 1e30)
    assert np.allclose(out[0, :, 0], x ** g, rtol=2e-6)
    assert np.all(out[..., 3] == 1.0)
    assert np.array_equal(out8[0, :, 0], np.rint(np.clip(x ** g, 0, 1) * 255).astype(np.uint8))
    # filmic: the white point 11.2 maps to 1 (calc(W) / calc(W)), black to 0 (post_proc.glsl:22-32)
    out, out8 = O.post_proc(film, 1.0, 1)
    # (black: calc(0) = D*E/(D*F) - E/F leaves an fp32 rounding residue of ~1e-9, which the gamma lifts to ~1e-4)
    assert 0.0 <= out[0, 0, 0] < 1e-3 and out8[0, 0, 0] == 0 and abs(out[0, 3, 0] - 1.0) < 1e-6 and out8[0, 3, 0] == 255
    def calc(v):
        A, B, C_, D, E, F = 0.22, 0.3, 0.1, 0.2, 0.01, 0.3
        return (v * (v * A + B * C_) + D * E) / (v * (v * A + B) + D * F) - E / F
    assert np.allclose(out[0, 1:3, 0], (calc(np.array([0.18, 1.0]) * 1.6) / calc(11.2)) ** g, rtol=1e-5)
    # ACES (post_proc.glsl:34-37)
    out, _ = O.post_proc(film, 1.0, 2)
    v = np.array([0.18, 1.0])
    assert np.allclose(out[0, 1:3, 0], ((v * (v * 2.51 + 0.03)) / (v * (v * 2.43 + 0.59) + 0.14)) ** g, rtol=1e-5)
    assert out[0, 4, 0] == 0.0                      # negative radiance clamps to 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind,tone", [("path", "filmic"), ("path", "aces"), ("path", "none"), ("light", "filmic")])
def test_post_process_matches_oracle(kind, tone, zl):
    from conftest import get_scene
    w, h = 64, 48
    s, _ = get_scene("cornell", w, h)
    s.upload()
    integ = (zl.NaivePathIntegrator if kind == "path" else zl.LightPathIntegrator)(s, w, h)
    for _ in range(4):
        integ.renderOnePass()
    film = integ.getFrame(1.0)                      # raw sums
    tm = integ.TONE_MAPPERS[tone]
    rgba, rgb8 = integ.postProcess(tone)
    ref, ref8 = O.post_proc(film, integ.trueScale(), tm)
    # same IEEE operations on both sides except powf (CUDA vs glibc: <= 2 ulp): relative 1e-6
    assert np.allclose(rgba, ref, rtol=1e-6, atol=1e-7, equal_nan=True)
    assert np.max(np.abs(rgb8.astype(np.int32) - ref8.astype(np.int32))) <= 1 and np.mean(rgb8 != ref8) < 1e-3
    assert rgb8.max() > 0
    # explicit scale, and the C ABI's argument checking
    rgba2, _ = integ.postProcess(tone, scale=0.25)
    ref2, _ = O.post_proc(film, 0.25, tm)
    assert np.allclose(rgba2, ref2, rtol=1e-6, atol=1e-7, equal_nan=True)
    with pytest.raises(zl.ZillumError):
        integ.postProcess(7)


@pytest.mark.gpu
def test_headless_cli_writes_the_screenshot(zl, tmp_path):
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(zl.__file__), "host", "zillum_render")
    png, pfm = tmp_path / "c.png", tmp_path / "c.pfm"
    r = subprocess.run([exe, "builtin:cornell", "--integrator", "path", "--spp", "4", "--size", "48x32", "--out", str(pfm),
                        "--png", str(png), "--tonemap", "aces"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    img = _read_png(png)
    assert img.shape == (32, 48, 3) and img.max() > 0
    from conftest import get_scene
    s, _ = get_scene("cornell", 48, 32)
    s.upload()
    integ = zl.NaivePathIntegrator(s, 48, 32)
    for _ in range(4):
        integ.renderOnePass()
    _, rgb8 = integ.postProcess("aces")
    assert np.array_equal(img, rgb8[::-1])
