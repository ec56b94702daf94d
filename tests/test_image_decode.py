"""Texture decoders of the scene-ingestion row (stbi_load(path, ..., 3) in src/core/Image.cpp:10-34):
PNG / JPEG (sequential and progressive) / TGA / BMP / PPM -> 8-bit RGB, checked against files written (and, for JPEG,
decoded) by Pillow."""
import struct
import zlib

import numpy as np
import pytest

PIL = pytest.importorskip("PIL.Image")


def _picture(w, h, seed=0):
    """Smooth colour ramps + a block of noise: exercises PNG filters, RLE runs and JPEG AC coefficients."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([(x * 255 // max(w - 1, 1)), (y * 255 // max(h - 1, 1)), ((x + y) * 255 // max(w + h - 2, 1))], axis=-1).astype(np.uint8)
    img[h // 4:h // 2, w // 4:w // 2] = rng.integers(0, 256, (h // 2 - h // 4, w // 2 - w // 4, 3), dtype=np.uint8)
    img[: h // 8, : w // 3] = (200, 30, 90)                       # a flat area (RLE runs)
    return img


@pytest.mark.parametrize("w,h", [(37, 23), (64, 64), (1, 1), (130, 5)])
def test_png_rgb_rgba_grey_palette_16bit(w, h, zl, tmp_path):
    img = _picture(w, h, seed=w)
    p = tmp_path / "a.png"
    PIL.fromarray(img).save(p)                                     # 8-bit RGB, adaptive filters, dynamic Huffman
    assert np.array_equal(zl.load_byte_image(p), img)
    PIL.fromarray(img).save(p, compress_level=0)                   # stored blocks
    assert np.array_equal(zl.load_byte_image(p), img)
    rgba = np.concatenate([img, np.full((h, w, 1), 77, np.uint8)], axis=-1)
    PIL.fromarray(rgba).save(p)                                    # alpha is dropped
    assert np.array_equal(zl.load_byte_image(p), img)
    grey = img[..., 0]
    PIL.fromarray(grey).save(p)                                    # grey is replicated
    assert np.array_equal(zl.load_byte_image(p), np.repeat(grey[..., None], 3, axis=-1))
    pal = PIL.fromarray(img).quantize(16)                          # 4-bit palette
    pal.save(p)
    assert np.array_equal(zl.load_byte_image(p), np.asarray(pal.convert("RGB")))
    g16 = (img[..., 1].astype(np.uint16) << 8) | 0x5a
    PIL.fromarray(g16).save(p)                                     # 16-bit grey: high byte
    assert np.array_equal(zl.load_byte_image(p)[..., 0], img[..., 1])
    bw = PIL.fromarray(((img[..., 0] > 127) * 255).astype(np.uint8)).convert("1")
    bw.save(p)                                                     # 1 bit per pixel
    assert np.array_equal(zl.load_byte_image(p)[..., 0], np.asarray(bw.convert("L")))


def _png_chunk(kind, data):
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xffffffff)


def test_png_adam7_interlaced(zl, tmp_path):
    """Pillow cannot write interlaced files: build one by hand (filter 0 and filter 1 rows alternate)."""
    w, h = 29, 19
    img = _picture(w, h, seed=5)
    x0, y0, dx, dy = [0, 4, 0, 2, 0, 1, 0], [0, 0, 4, 0, 2, 0, 1], [8, 8, 4, 4, 2, 2, 1], [8, 8, 8, 4, 4, 2, 2]
    raw = bytearray()
    for p in range(7):
        sub = img[y0[p]::dy[p], x0[p]::dx[p]]
        for r, row in enumerate(sub):
            flat = row.reshape(-1).astype(np.int32)
            if r % 2 == 0:
                raw += b"\x00" + flat.astype(np.uint8).tobytes()
            else:                                                  # Sub filter, bpp = 3
                left = np.concatenate([np.zeros(3, np.int32), flat[:-3]])
                raw += b"\x01" + ((flat - left) & 255).astype(np.uint8).tobytes()
    f = tmp_path / "i.png"
    f.write_bytes(b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 1))
                  + _png_chunk(b"IDAT", zlib.compress(bytes(raw), 9)) + _png_chunk(b"IEND", b""))
    assert np.array_equal(np.asarray(PIL.open(f).convert("RGB")), img)      # the hand-made file is valid
    assert np.array_equal(zl.load_byte_image(f), img)


@pytest.mark.parametrize("kw", [dict(), dict(compression="tga_rle")])
def test_tga_and_bmp(kw, zl, tmp_path):
    img = _picture(53, 31, seed=2)
    p = tmp_path / "a.tga"
    PIL.fromarray(img).save(p, **kw)
    assert np.array_equal(zl.load_byte_image(p), img)
    rgba = np.concatenate([img, np.full(img.shape[:2] + (1,), 200, np.uint8)], axis=-1)
    PIL.fromarray(rgba).save(p, **kw)
    assert np.array_equal(zl.load_byte_image(p), img)
    PIL.fromarray(img[..., 2]).save(p, **kw)
    assert np.array_equal(zl.load_byte_image(p)[..., 1], img[..., 2])
    b = tmp_path / "a.bmp"
    PIL.fromarray(img).save(b)                                     # 24-bit rows padded to 4 bytes (53 * 3 = 159)
    assert np.array_equal(zl.load_byte_image(b), img)
    PIL.fromarray(img).quantize(64).save(b)                        # 8-bit palette
    assert np.array_equal(zl.load_byte_image(b), np.asarray(PIL.open(b).convert("RGB")))


@pytest.mark.parametrize("kw", [dict(quality=92, subsampling=0), dict(quality=85, subsampling=2), dict(quality=75, subsampling=1),
                                dict(quality=95, subsampling=0, optimize=True), dict(quality=90, grey=True),
                                dict(quality=88, subsampling=2, restart_marker_blocks=3),
                                dict(quality=90, subsampling=2, progressive=True), dict(quality=80, subsampling=0, progressive=True, optimize=True),
                                dict(quality=85, grey=True, progressive=True), dict(quality=75, subsampling=1, progressive=True)])
def test_jpeg_close_to_pillow(kw, zl, tmp_path):
    """JPEG decoders differ in IDCT rounding and chroma up-sampling (stb's does too), so the check is a tolerance
    against Pillow's libjpeg: a few grey levels at most, well under one level on average."""
    kw = dict(kw)
    img = _picture(75, 46, seed=9)                                 # not a multiple of the MCU size
    src = PIL.fromarray(img[..., 0] if kw.pop("grey", False) else img)
    p = tmp_path / "a.jpg"
    try:
        src.save(p, **kw)
    except TypeError:
        pytest.skip("this Pillow cannot write restart markers")
    ref = np.asarray(PIL.open(p).convert("RGB")).astype(np.int32)
    got = zl.load_byte_image(p)
    assert got is not None and got.shape == ref.shape
    d = np.abs(got.astype(np.int32) - ref)
    assert d.mean() < 0.75 and np.percentile(d, 99.5) <= 4, (d.mean(), d.max())
    assert np.abs(got.astype(np.int32) - np.asarray(src.convert("RGB"))).mean() < 12      # and it IS the picture


def test_unsupported_and_broken_files_are_refused(zl, tmp_path):
    img = _picture(40, 40)
    p = tmp_path / "p.jpg"
    PIL.fromarray(img).convert("CMYK").save(p)
    assert zl.load_byte_image(p) is None                           # 4-component JPEG: refused, not mis-decoded
    PIL.fromarray(img).save(p, progressive=True)
    data = p.read_bytes()
    p.write_bytes(data[: len(data) * 2 // 3])
    zl.load_byte_image(p)                                          # a truncated progressive file: any answer, no crash
    q = tmp_path / "t.png"
    PIL.fromarray(img).save(q)
    data = q.read_bytes()
    q.write_bytes(data[: len(data) // 2])
    assert zl.load_byte_image(q) is None                           # truncated
    assert zl.load_byte_image(tmp_path / "missing.png") is None
    (tmp_path / "e.tga").write_bytes(b"")
    assert zl.load_byte_image(tmp_path / "e.tga") is None


def test_obj_material_with_png_texture(zl, tmp_path):
    """map_Kd may name any supported format: the model importer goes through the same loader."""
    tex = _picture(16, 16, seed=3)
    PIL.fromarray(tex).save(tmp_path / "albedo.png")
    (tmp_path / "m.mtl").write_text("newmtl m0\nKd 1 1 1\nmap_Kd albedo.png\n")
    (tmp_path / "m.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nvn 0 0 1\nusemtl m0\nf 1/1/1 2/2/1 3/3/1\n")
    (tmp_path / "scene.xml").write_text(
        '<?xml version="1.0"?>\n<scene name="t">\n<integrator type="path"><maxBounce value="3" /><size width="16" height="16" /></integrator>\n'
        '<sampler type="sobol"><numSamples value="4" /></sampler>\n'
        '<camera type="thinLens"><position value="0 -3 0" /><angle value="0 0 0" /><fov value="45" /><lensRadius value="0" /><focalDistance value="1" /></camera>\n'
        f'<modelInstances><modelInstance path="{tmp_path / "m.obj"}" name="tri" type="object">'
        '<transform translate="0 0 0" scale="1 1 1" rotate="0 0 0" /><material type="default" /></modelInstance></modelInstances>\n</scene>\n')
    s = zl.Scene.from_file(tmp_path / "scene.xml")
    s.flatten()
    assert s.info["numTextures"] == 1 and s.info["numTriangles"] == 1


def test_ldr_environment_map_is_converted_like_stbi_loadf(zl, tmp_path):
    """The commented res/scene.xml names a .png environment map; the reference reads it with stbi_loadf
    (src/core/EnvironmentMap.cpp:8-15 via Image.cpp:16-19), i.e. pow(v / 255, 2.2) per channel."""
    img = _picture(32, 16, seed=11)
    PIL.fromarray(img).save(tmp_path / "sky.png")
    (tmp_path / "scene.xml").write_text(
        '<?xml version="1.0"?>\n<scene name="t">\n<integrator type="path"><maxBounce value="3" /><size width="16" height="16" /></integrator>\n'
        '<sampler type="sobol"><numSamples value="4" /></sampler>\n'
        '<camera type="thinLens"><position value="0 -3 0" /><angle value="0 0 0" /><fov value="45" /><lensRadius value="0" /><focalDistance value="1" /></camera>\n'
        f'<modelInstances></modelInstances>\n<envMap path="{tmp_path / "sky.png"}" />\n</scene>\n')
    s = zl.Scene.from_file(tmp_path / "scene.xml")
    info = s.info
    assert (info["envW"], info["envH"]) == (32, 16)
    with pytest.raises(zl.ZillumError):          # no triangles (the reference would crash in BVH::build): reported, not fatal
        s.flatten()
    env = s.array("envMap")
    want = np.power((img.astype(np.float32) / np.float32(255.0)).astype(np.float64), 2.2).astype(np.float32)
    assert np.array_equal(env.reshape(16, 32, 3), want)


def test_png_remaining_colour_types(zl, tmp_path):
    """grey + alpha (type 4), 16-bit RGB / RGBA (types 2 / 6 at depth 16), 2-bit palette — written by hand where Pillow has no writer."""
    h, w = 9, 13
    img = _picture(w, h, seed=21)
    la = np.stack([img[..., 0], np.full((h, w), 99, np.uint8)], axis=-1)
    p = tmp_path / "la.png"
    PIL.fromarray(la, "LA").save(p)
    assert np.array_equal(zl.load_byte_image(p), np.repeat(img[..., :1], 3, axis=-1))

    def write(path, ctype, depth, rows):
        raw = b"".join(b"\x00" + r for r in rows)
        path.write_bytes(b"\x89PNG\r\n\x1a\n" + _png_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
                         + _png_chunk(b"IDAT", zlib.compress(raw, 6)) + _png_chunk(b"IEND", b""))
    rgb16 = (img.astype(np.uint16) << 8) | 0x33                       # big-endian samples; the loader keeps the high byte
    write(tmp_path / "rgb16.png", 2, 16, [rgb16[y].astype(">u2").tobytes() for y in range(h)])
    assert np.array_equal(zl.load_byte_image(tmp_path / "rgb16.png"), img)
    rgba16 = np.concatenate([rgb16, np.full((h, w, 1), 0xabcd, np.uint16)], axis=-1)
    write(tmp_path / "rgba16.png", 6, 16, [rgba16[y].astype(">u2").tobytes() for y in range(h)])
    assert np.array_equal(zl.load_byte_image(tmp_path / "rgba16.png"), img)
    assert np.array_equal(np.asarray(PIL.open(tmp_path / "rgb16.png").convert("RGB")), img)        # the hand-made files are valid
    g2 = (img[..., 1] >> 6).astype(np.uint8)                          # 2-bit grey: 0..3 -> 0, 85, 170, 255
    rows = []
    for y in range(h):
        bits = np.zeros(((w + 3) // 4) * 4, np.uint8); bits[:w] = g2[y]
        rows.append(bytes((bits[0::4] << 6) | (bits[1::4] << 4) | (bits[2::4] << 2) | bits[3::4]))
    write(tmp_path / "g2.png", 0, 2, rows)
    assert np.array_equal(zl.load_byte_image(tmp_path / "g2.png")[..., 2], g2 * 85)


def _scene_xml(tmp_path, obj):
    (tmp_path / "scene.xml").write_text(
        '<?xml version="1.0"?>\n<scene name="t">\n<integrator type="path"><maxBounce value="3" /><size width="16" height="16" /></integrator>\n'
        '<sampler type="sobol"><numSamples value="4" /></sampler>\n'
        '<camera type="thinLens"><position value="0 -3 0" /><angle value="0 0 0" /><fov value="45" /><lensRadius value="0" /><focalDistance value="1" /></camera>\n'
        f'<modelInstances><modelInstance path="{obj}" name="m" type="object">'
        '<transform translate="0 0 0" scale="1 1 1" rotate="0 0 0" /><material type="default" /></modelInstance></modelInstances>\n</scene>\n')
    return tmp_path / "scene.xml"


def test_obj_negative_indices_polygons_and_windows_style_mtl(zl, tmp_path):
    """Relative (negative) indices count back from the elements read so far, so the same token means different vertices
    in different faces; polygons become fans; a Windows-authored MTL names its texture with options and backslashes."""
    (tmp_path / "textures").mkdir()
    PIL.fromarray(_picture(8, 8, seed=4)).save(tmp_path / "textures" / "wall.tga")
    (tmp_path / "w.mtl").write_text("newmtl wall\r\nKd 0.5 0.5 0.5\r\nmap_Kd -s 1 1 1 textures\\wall.tga\r\n")
    (tmp_path / "w.obj").write_text(
        "mtllib w.mtl\r\nusemtl wall\r\n"
        "v 0 0 0\r\nv 1 0 0\r\nv 1 1 0\r\nv 0 1 0\r\nvt 0 0\r\nvt 1 0\r\nvt 1 1\r\nvt 0 1\r\n"
        "f -4/-4 -3/-3 -2/-2 -1/-1\r\n"                      # a quad through relative indices -> two triangles
        "v 0 0 1\r\nv 1 0 1\r\nv 1 1 1\r\nvt 0 0\r\nvt 1 0\r\nvt 1 1\r\n"
        "f -3/-3 -2/-2 -1/-1\r\n"                            # the same tokens, other vertices
        "f 1/1 2/2 99/1\r\n")                                # out-of-range index: skipped with a message, not a crash
    s = zl.Scene.from_file(_scene_xml(tmp_path, tmp_path / "w.obj"))
    s.flatten()
    assert s.info["numTriangles"] == 3 and s.info["numTextures"] == 1
    v = s.array("vertices").reshape(-1, 3)
    idx = s.array("indices").reshape(-1, 3)
    used = {tuple(np.round(v[i], 5)) for i in idx.reshape(-1)}
    assert len(used) == 7                                    # 4 quad corners + 3 other vertices: the repeated tokens were not joined
