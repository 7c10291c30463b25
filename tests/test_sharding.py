"""Multi-rank logic on CPU: two gloo processes shard a batch of independent scenes, render their
share with the CPU oracle, and the gathered per-scene checksums must equal an unsharded render; the timing
reduction must be max-over-ranks / sum-of-units."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from z2d_b200 import abi, sharding, workloads

N_SCENES, SIZE, PATHS = 6, 96, 24


def _render_scene(scene_index):
    from tests.oracle_backend import load_oracle
    lib = load_oracle()
    scene = workloads.cubic_paths_scene(PATHS, SIZE, seed=sharding.scene_seed(sharding.BASE_SEED_C5, scene_index), r_log2=(2.0, 5.0))
    buf = np.zeros(SIZE * SIZE * 4, dtype=np.uint8)
    cmds = scene.draw_cmds(0)
    P = C.POINTER
    for i in range(scene.n):
        rc = lib.z2d_ref_fill(buf.ctypes.data_as(C.c_void_p), int(abi.Format.rgba), SIZE, SIZE,
                              C.cast(C.c_void_p(int(cmds["pattern"][i])), P(abi.PatternPOD)),
                              C.cast(C.c_void_p(int(cmds["nodes"][i])), P(abi.Node)), int(cmds["n_nodes"][i]),
                              C.cast(C.c_void_p(int(cmds["fill"][i])), P(abi.FillOptsPOD)))
        assert rc == 0
    return buf


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.scenes_of_rank(N_SCENES, world, rank)
        local = {s: sharding.surface_checksum(_render_scene(s)) for s in mine}
        sums = sharding.gather_checksums(local, N_SCENES, dist)
        ms, units = sharding.reduce_timing([10.0 + rank, 5.0 - rank], [len(mine), 100.0 * (rank + 1)], dist=dist)
        np.save(os.path.join(out_dir, f"rank{rank}.npy"),
                np.array([b"".join(sums).hex(), repr(ms), repr(units), repr(mine)], dtype=object), allow_pickle=True)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_scene_assignment_partitions_the_batch():
    for world in (1, 2, 4, 8):
        seen = sorted(s for r in range(world) for s in sharding.scenes_of_rank(4096, world, r))
        assert seen == list(range(4096))
        sizes = [len(sharding.scenes_of_rank(4096, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.scenes_of_rank(8, 2, 2)


def test_gather_detects_missing_and_duplicate_scenes():
    with pytest.raises(RuntimeError):
        sharding.gather_checksums({0: b"x" * 32}, 2)


@pytest.mark.timeout(300)
def test_two_rank_sharded_render_equals_unsharded(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    expect = b"".join(sharding.surface_checksum(_render_scene(s)) for s in range(N_SCENES)).hex()
    owned = []
    for r in range(world):
        sums, ms, units, mine = np.load(os.path.join(tmp_path, f"rank{r}.npy"), allow_pickle=True)
        assert sums == expect, f"rank {r}: gathered checksums differ from the unsharded render"
        assert eval(ms) == [11.0, 5.0]            # max over ranks
        assert eval(units) == [float(N_SCENES), 300.0]  # sum over ranks
        owned += eval(mine)
    assert sorted(owned) == list(range(N_SCENES))
