"""Multi-GPU partition logic on CPU: world_size 2, gloo backend.

Each rank drives the real C++ Integrator bookkeeping (setSampleShard: rank g of G renders passes
g, g+G, ... with the uSpp / uFreeCounter a single GPU would use), renders its passes with the
CPU ORACLE in place of the device kernels (this is the checker standing in for the GPU, test
only), and the films are summed with ONE all-reduce -- the only data-path collective.  The sum
must equal the film of a single process rendering all passes."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT

W, H, PASSES, WORLD = 24, 18, 6, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _render_shard(rank, world, kind):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_lib as O
    import zillumgl_b200 as zl
    O.lib.zo_set_threads(1)
    scene = zl.Scene.builtin("cornell", W, H)
    scene.flatten()
    oracle = O.OracleScene(scene.desc)
    cls = {"path": zl.NaivePathIntegrator, "triple": zl.TriplePathIntegrator}[kind]
    integ = cls(scene, W, H, host_only=True)
    if kind == "triple":
        integ.mParam.LPTBlocksOnePass = 1
    integ.setSampleShard(rank, world)
    film = np.zeros((H, W, 4), np.float32)
    seen = []
    for _ in range(PASSES // world):
        p = integ.params(0)
        seen.append((p.spp, p.freeCounter))
        if kind == "path":
            oracle.path_pass(p, film)
        else:
            oracle.triple_pt_pass(p, film)
            oracle.triple_lpt_pass(integ.params(1), film)
        integ.renderOnePass()              # host bookkeeping only (no device): advances the pass index by `world`
    return film, seen


def _worker(rank, world, port, kind, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    film, seen = _render_shard(rank, world, kind)
    t = torch.from_numpy(film)
    dist.all_reduce(t)                     # the film sum: the one collective of the path
    idx = torch.tensor(seen, dtype=torch.int64)
    gathered = [torch.zeros_like(idx) for _ in range(world)]
    dist.all_gather(gathered, idx)
    if rank == 0:
        np.save(os.path.join(out_dir, "film.npy"), t.numpy())
        np.save(os.path.join(out_dir, "passes.npy"), torch.stack(gathered).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["path", "triple"])
def test_sample_sharding_world_size_2_gloo(kind, tmp_path):
    mp = pytest.importorskip("torch.multiprocessing")
    mp.spawn(_worker, args=(WORLD, _free_port(), kind, str(tmp_path)), nprocs=WORLD, join=True)
    film = np.load(tmp_path / "film.npy")
    passes = np.load(tmp_path / "passes.npy")          # [rank][k] = (spp, freeCounter)
    # every pass index 0..PASSES-1 rendered exactly once, with the single-GPU free counter
    assert sorted(passes[..., 0].ravel().tolist()) == list(range(PASSES))
    assert np.array_equal(passes[..., 1], passes[..., 0] + 1)
    assert np.array_equal(passes[1, :, 0], passes[0, :, 0] + 1)
    whole, seen = _render_shard(0, 1, kind)
    # the unsharded integrator counts its free counter up from 1 as well
    assert [s for s, _ in seen] == list(range(PASSES)) and [f for _, f in seen] == list(range(1, PASSES + 1))
    assert whole[..., :3].max() > 0
    assert np.allclose(film, whole, rtol=1e-5, atol=1e-6)   # equal up to FP32 summation order


def _worker_progressive(rank, world, port, out_dir):
    """bench.py's end-to-end schedule at N > 1 (DESIGN.md §6 "reduce before copy"): every step each rank renders one pass, a copy of its
    film is reduce-scattered, rank r keeps rows [r*H/N, (r+1)*H/N) of the sum = the progressive frame of all passes so far."""
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_lib as O
    import zillumgl_b200 as zl
    O.lib.zo_set_threads(1)
    scene = zl.Scene.builtin("cornell", W, H)
    scene.flatten()
    oracle = O.OracleScene(scene.desc)
    integ = zl.NaivePathIntegrator(scene, W, H, host_only=True)
    integ.setSampleShard(rank, world)
    film = np.zeros((H, W, 4), np.float32)
    hs = H // world
    rows = []
    for _ in range(PASSES // world):
        oracle.path_pass(integ.params(0), film)
        integ.renderOnePass()
        snap = torch.from_numpy(film.copy())                 # Integrator.snapshotAsync: a consistent copy of this rank's film, in pass order
        part = torch.empty((hs, W, 4), dtype=torch.float32)
        dist.reduce_scatter_tensor(part, snap)
        rows.append(part.numpy().copy())
    np.save(os.path.join(out_dir, f"rows{rank}.npy"), np.stack(rows))
    dist.barrier()
    dist.destroy_process_group()


def test_progressive_frames_by_reduce_scatter_world_size_2_gloo(tmp_path):
    mp = pytest.importorskip("torch.multiprocessing")
    assert H % WORLD == 0
    mp.spawn(_worker_progressive, args=(WORLD, _free_port(), str(tmp_path)), nprocs=WORLD, join=True)
    got = np.concatenate([np.load(tmp_path / f"rows{r}.npy") for r in range(WORLD)], axis=1)      # [step][H][W][4]
    # a single process rendering every pass in order: after step k the distributed frame holds passes 0 .. WORLD*(k+1)-1
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_lib as O
    import zillumgl_b200 as zl
    scene = zl.Scene.builtin("cornell", W, H)
    scene.flatten()
    oracle = O.OracleScene(scene.desc)
    integ = zl.NaivePathIntegrator(scene, W, H, host_only=True)
    film = np.zeros((H, W, 4), np.float32)
    for k in range(PASSES):
        oracle.path_pass(integ.params(0), film)
        integ.renderOnePass()
        if (k + 1) % WORLD == 0:
            step = (k + 1) // WORLD - 1
            assert np.allclose(got[step], film, rtol=1e-5, atol=1e-6), f"frame after step {step}"
    assert film[..., :3].max() > 0
