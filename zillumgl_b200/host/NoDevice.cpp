// Host-preparation-only build (libzillum_hostprep.so): the C++ host classes WITHOUT the CUDA library.  Every C-ABI entry point the
// host calls is defined here to fail with ZL_ERR_NO_DEVICE, so scene loading / flattening (procedural meshes, BVH, tables) works and
// anything that would render fails loudly.  Used by `bench.py --impl reference` only, whose CPU arm must not load libzillum_cuda.so;
// the product (libzillum_host.so) links the real library instead of this file.
#include "../../include/zillum_cuda.h"

extern "C" {
static int none() { return ZL_ERR_NO_DEVICE; }
const char* zl_last_error_string(void) { return "host-preparation-only build: libzillum_cuda.so is not linked"; }
int zl_set_device(int) { return none(); }
int zl_device_synchronize(void) { return none(); }
int zl_scene_create(const ZlSceneDesc*, ZlScene**) { return none(); }
int zl_scene_destroy(ZlScene*) { return none(); }
int zl_scene_prep_times(const ZlScene*, double*, double*, int*) { return none(); }
int zl_film_create(int, int, ZlFilm**) { return none(); }
int zl_film_create_external(int, int, void*, ZlFilm**) { return none(); }
int zl_film_destroy(ZlFilm*) { return none(); }
int zl_film_clear(ZlFilm*, void*) { return none(); }
int zl_film_flush(ZlFilm*, void*) { return none(); }
int zl_film_download(ZlFilm*, float, float*, void*) { return none(); }
int zl_film_download_async(ZlFilm*, float, float*, void*) { return none(); }
int zl_film_download_rgb_async(ZlFilm*, float, float*, void*) { return none(); }
int zl_film_download_wait(ZlFilm*) { return none(); }
int zl_film_snapshot_async(ZlFilm*, void*, void*) { return none(); }
int zl_film_postprocess(ZlFilm*, float, int, float*, unsigned char*, void*) { return none(); }
int zl_launch_path_pass(ZlScene*, ZlFilm*, const ZlRenderParams*, int, void*) { return none(); }
int zl_launch_light_pass(ZlScene*, ZlFilm*, const ZlRenderParams*, int, void*) { return none(); }
int zl_launch_triple_pt_pass(ZlScene*, ZlFilm*, const ZlRenderParams*, int, void*) { return none(); }
int zl_launch_triple_lpt_pass(ZlScene*, ZlFilm*, const ZlRenderParams*, int, void*) { return none(); }
}
