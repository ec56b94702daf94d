"""zl_scene_update_materials — the C-ABI stand-in for the material editor's `glContext.material->write(...)`
(src/gui/Editor.cpp:73): records rewritten in place on an uploaded scene must behave exactly like a scene that was
created with them."""
import ctypes as C

import numpy as np
import pytest

from conftest import get_scene


def test_update_materials_refuses_bad_arguments_without_touching_memory(zl):
    """argument checks come before any device work: a null scene or null records are ZL_ERR_INVALID_ARGUMENT (CPU-runnable)"""
    from zillumgl_b200 import _native as N
    m = np.zeros(16, np.float32)
    assert N.cuda.zl_scene_update_materials(None, 0, 1, m.ctypes.data_as(C.POINTER(C.c_float))) != 0
    assert b"zl_scene_update_materials" in N.cuda.zl_last_error_string()
    s = zl.Scene.builtin("cornell", 32, 24)
    s.flatten()
    with pytest.raises(zl.ZillumError):
        s.update_materials(0, m)                      # not uploaded


@pytest.mark.gpu
def test_updated_materials_equal_a_scene_created_with_them(zl):
    import ref_lib
    from test_ref_parity import assert_same_bits, bsdf_kat_inputs, params, random_materials, scene_with_materials
    rng = np.random.default_rng(900)
    w, h = 48, 27
    probe, _ = get_scene("sponza_light", w, h)
    count = probe.info["numMaterials"]
    mats = random_materials(rng, count)
    direct = scene_with_materials(zl, "sponza_light", w, h, mats)
    direct.upload()
    edited = zl.Scene.builtin("sponza_light", w, h)
    edited.flatten()
    edited.upload()
    edited.update_materials(0, mats[: count // 2])                 # two partial writes, like two edits in the GUI
    edited.update_materials(count // 2, mats[count // 2:])
    with pytest.raises(zl.ZillumError):
        edited.update_materials(count - 1, mats[:2])               # past the end of the table
    p = params(zl, direct, w, h)
    for op, inp, nout in bsdf_kat_inputs(rng, 1024, range(count)):
        assert_same_bits(zl.debug_eval(edited, p, zl.KAT[op], inp, nout), zl.debug_eval(direct, p, zl.KAT[op], inp, nout), (op, int(inp[0, 0].view(np.int32))))
    ref = ref_lib.RefScene(direct.desc) if ref_lib.available() else None
    for variant in (0, 2):
        frames = []
        for scene in (edited, direct):
            integ = zl.NaivePathIntegrator(scene, w, h)
            integ.mParam.kernelVariant = variant
            film = np.zeros((h, w, 4), np.float32)
            for _ in range(3):
                if ref is not None and scene is direct:
                    ref.path_pass(integ.params(), film)
                integ.renderOnePass()
            frames.append(np.ascontiguousarray(integ.getFrame(1.0)[..., :3]))
        assert_same_bits(frames[0], frames[1], ("film", variant))
        if ref is not None:
            assert_same_bits(frames[1], np.ascontiguousarray(film[..., :3]), ("film vs the reference's shaders", variant))
