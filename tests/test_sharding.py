"""Host-side logic of the multi-GPU path on CPU: partitioning, sub-batching and
the final result gather with torch.distributed (gloo, world_size 2)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from vanishing_points_2017_b200 import pipeline, synth


def test_shards_partition_the_batch_and_balance_cost():
    n, _ = synth.config_sizes(4, 400)
    off = np.concatenate([[0], np.cumsum(n)])
    for ws in (1, 2, 4, 8):
        parts = [pipeline.shard_batch(off, ws, r) for r in range(ws)]
        allidx = np.sort(np.concatenate(parts))
        np.testing.assert_array_equal(allidx, np.arange(400))
        loads = np.array([pipeline.image_cost(n[p]).sum() for p in parts])
        assert loads.max() / loads.mean() < 1.02


def test_take_images_roundtrip():
    b = synth.make_batch(2, n_images=6)
    idx = np.array([4, 1, 5])
    seg, off = pipeline.take_images(b["segments"], b["offsets"], idx)
    for k, i in enumerate(idx):
        np.testing.assert_array_equal(seg[off[k]:off[k + 1]], b["segments"][b["offsets"][i]:b["offsets"][i + 1]])
    seg0, off0 = pipeline.take_images(b["segments"], b["offsets"], np.array([], dtype=int))
    assert seg0.shape == (0, 4) and off0.tolist() == [0]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_images = 11
    off = np.concatenate([[0], np.cumsum(np.arange(100, 100 + n_images))])
    idx = pipeline.shard_batch(off, world, rank)
    local = [{"vp": np.full((1, 3), float(i)), "rank": rank} for i in idx]
    out = pipeline.gather_results(local, idx, n_images, world, dist)
    if rank == 0:
        q.put([(int(r["vp"][0, 0]), r["rank"]) for r in out])
    dist.barrier()
    dist.destroy_process_group()


def test_gather_results_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [g[0] for g in got] == list(range(11))
    assert {g[1] for g in got} == {0, 1}


def _fake_raw(idx, n):
    """Flat result arrays as vp_localisation._alloc_result lays them out, filled with values that encode
    the global image index."""
    B, M = len(idx), 64
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int32)
    arrs = {"status": np.zeros(B, np.int32), "n_vp": (np.asarray(idx) % 5 + 1).astype(np.int32),
            "iterations": np.asarray(idx, np.int32) * 2, "vp": np.zeros((B, M, 3)), "sigma": np.zeros((B, M)),
            "counts": np.zeros((B, M), np.int32), "counts_weighted": np.zeros((B, M)),
            "vp_assoc": np.full(max(int(off[-1]), 1) + 7, -5, np.int32)}       # longer than sum N, like a reused buffer
    for k, i in enumerate(idx):
        arrs["vp"][k, 0] = [i, i + 0.5, -i]
        arrs["vp_assoc"][off[k]:off[k + 1]] = i
    return arrs, off


def _worker_raw(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_all = np.array([30, 0, 75, 12, 51, 9, 64])
    off_all = np.concatenate([[0], np.cumsum(n_all)])
    idx = pipeline.shard_batch(off_all, world, rank)
    arrs, off = _fake_raw(idx, n_all[idx])
    out = pipeline.gather_raw(arrs, off_all, world, rank, dist)
    if rank == 0:
        q.put({k: v for k, v in out.items()})
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_raw_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_raw, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_all = np.array([30, 0, 75, 12, 51, 9, 64])
    np.testing.assert_array_equal(out["n"], n_all)
    np.testing.assert_array_equal(out["iterations"], np.arange(7) * 2)
    np.testing.assert_array_equal(out["vp"][:, 0, 0], np.arange(7.0))
    assert out["vp_assoc"].shape[0] == n_all.sum()
    for i in range(7):              # image i owns vp_assoc[assoc_start[i] : assoc_start[i] + n[i]] (rank-major order)
        a = out["assoc_start"][i]
        np.testing.assert_array_equal(out["vp_assoc"][a:a + n_all[i]], np.full(n_all[i], i))
    np.testing.assert_array_equal(out["n_vp"], np.arange(7) % 5 + 1)
    # single rank: the identity
    arrs, off = _fake_raw(np.array([0, 1, 2]), np.array([3, 4, 5]))
    one = pipeline.gather_raw(arrs, off, 1, 0)
    np.testing.assert_array_equal(one["vp_assoc"], np.repeat([0, 1, 2], [3, 4, 5]))
    np.testing.assert_array_equal(one["assoc_start"], [0, 3, 7])
    np.testing.assert_array_equal(one["vp"][:, 0, 1], np.arange(3) + 0.5)
