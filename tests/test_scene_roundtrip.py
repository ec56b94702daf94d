"""Scene ingestion round trip (reference: Resource.cpp:37-168 model import, Model.cpp:62-72 TRS order, Scene.cpp:58-127).
The procedural Sponza-class scene (262,144 triangles, 11 meshes, one textured) is EXPORTED to the formats the reference ingests —
Wavefront OBJ + MTL + a PNG albedo texture + scene.xml — re-imported through the product's own readers (host/Model.cpp,
host/ImageDecode.cpp, host/Xml.cpp, host/Scene.cpp) and flattened: geometry, BVH, MTBVH table, texture and colours must come back
bit for bit; on the GPU the re-imported scene must render the film the reference's shaders render for it."""
import os

import numpy as np
import pytest

PIL = pytest.importorskip("PIL.Image")


def _f(x):
    return np.format_float_positional(np.float32(x), unique=True, trim="-")      # shortest text that parses back to the same binary32


def export_obj(models, images, directory, name):
    """one OBJ + MTL for the OBJECT model instances of a scene; one material group per mesh (the reader forms one mesh per group)"""
    obj_path = os.path.join(directory, name + ".obj")
    with open(obj_path, "w") as o, open(os.path.join(directory, name + ".mtl"), "w") as m:
        o.write(f"mtllib {name}.mtl\n")
        base = 0
        k = 0
        for model in models:
            for mesh in model["meshes"]:
                mat = model["materials"][mesh["matIndex"]]
                m.write(f"newmtl mesh{k}\nKd {_f(mat[0])} {_f(mat[1])} {_f(mat[2])}\n")
                if mesh["texIndex"] >= 0:
                    tex = f"{name}_tex{mesh['texIndex']}.png"
                    PIL.fromarray(images[mesh["texIndex"]]).save(os.path.join(directory, tex))
                    m.write(f"map_Kd {tex}\n")
                o.write(f"usemtl mesh{k}\n")
                for p in mesh["pos"]:
                    o.write(f"v {_f(p[0])} {_f(p[1])} {_f(p[2])}\n")
                for t in mesh["tex"]:
                    o.write(f"vt {_f(t[0])} {_f(np.float32(1) - t[1])}\n")      # the reader flips v back (aiProcess_FlipUVs)
                for n in mesh["nrm"]:
                    o.write(f"vn {_f(n[0])} {_f(n[1])} {_f(n[2])}\n")
                idx = mesh["idx"].reshape(-1, 3).astype(np.int64) + base + 1
                o.write("".join(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}\n" for a, b, c in idx))
                base += mesh["pos"].shape[0]
                k += 1
    return obj_path


@pytest.fixture(scope="module")
def roundtrip(tmp_path_factory):
    import zillumgl_b200 as zl
    w, h = 64, 36
    d = str(tmp_path_factory.mktemp("sponza_export"))
    src = zl.Scene.builtin("sponza", w, h)
    src.flatten()
    models = src.models()
    objects = [m for m in models if not m["isLight"]]
    assert len(objects) == 1 and sum(x["idx"].size for x in objects[0]["meshes"]) // 3 == 262144
    obj = export_obj(objects, zl.Scene.images(), d, "sponza")
    env = src.array("envMap").reshape(src.desc.contents.envH, src.desc.contents.envW, 3)
    with open(os.path.join(d, "sky.pfm"), "wb") as f:                               # PFM: rows bottom-up, little endian
        f.write(b"PF\n%d %d\n-1.0\n" % (env.shape[1], env.shape[0]))
        f.write(np.ascontiguousarray(env[::-1], "<f4").tobytes())
    xml = src.builtin_xml("sponza", w, h)
    t = objects[0]["trs"]
    xml = xml.replace('path="builtin:sponza"', f'path="{obj}"').replace('path="builtin:sky"', f'path="{os.path.join(d, "sky.pfm")}"')
    assert obj in xml and "sky.pfm" in xml and tuple(t[:3]) == (0, 0, 0)
    path = os.path.join(d, "scene.xml")
    open(path, "w").write(xml)
    back = zl.Scene.from_file(path)
    back.flatten()
    return zl, src, back, w, h


def test_reimported_scene_flattens_to_the_same_arrays(roundtrip):
    zl, src, back, w, h = roundtrip
    for k in ("numTriangles", "bvhSize", "objPrimCount", "nLightTriangles", "numTextures", "envW", "envH"):
        assert back.info[k] == src.info[k], k
    for a in ("bounds", "hitTable", "texels", "texUVScale", "envAlias", "envAliasProb", "lightPower"):
        x, y = back.array(a), src.array(a)
        assert x.shape == y.shape and np.array_equal(x.view(np.uint8), y.view(np.uint8)), a
    # The reader numbers vertices in the order the faces first use them (and drops unused ones), like an importer that joins identical
    # vertices: the vertex ARRAYS are a permutation of the exported ones, the TRIANGLES (what the kernels see) are the same, in order.
    ib, is_ = back.array("indices"), src.array("indices")
    objv = 3 * src.info["objPrimCount"]
    for a, c in (("vertices", 3), ("normals", 3)):
        x, y = back.array(a).reshape(-1, c)[ib], src.array(a).reshape(-1, c)[is_]
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), a
    # v is flipped twice (writer 1 - v, reader 1 - v): equal to 1 - (1 - v) in binary32
    tx, ty = back.array("texcoords").reshape(-1, 2)[ib[:objv]], src.array("texcoords").reshape(-1, 2)[is_[:objv]]
    assert np.array_equal(tx[:, 0], ty[:, 0]) and np.array_equal(tx[:, 1], np.float32(1) - (np.float32(1) - ty[:, 1]))
    # MTL carries the diffuse colour only: per triangle the same base colour and the same texture layer
    mb, ms = back.array("materials").reshape(-1, 16), src.array("materials").reshape(-1, 16)
    ib, is_ = back.array("matTexIndices"), src.array("matTexIndices")
    assert np.array_equal(ib >> 16, is_ >> 16)
    assert np.array_equal(mb[ib & 0xffff, :3], ms[is_ & 0xffff, :3])
    assert np.array_equal(back.array("envMap"), src.array("envMap"))


@pytest.mark.gpu
def test_reimported_scene_renders_the_reference_film(roundtrip):
    import ref_lib
    if not ref_lib.available():
        pytest.skip("oracle/_ref not present")
    zl, src, back, w, h = roundtrip
    back.upload()
    ref = ref_lib.RefScene(back.desc)
    integ = zl.NaivePathIntegrator(back, w, h)
    film = np.zeros((h, w, 4), np.float32)
    for _ in range(4):
        ref.path_pass(integ.params(), film)
        integ.renderOnePass()
    got = np.ascontiguousarray(integ.getFrame(1.0)[..., :3])
    assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(film[..., :3]).view(np.uint32))
    assert film[..., :3].mean() > 1e-3
