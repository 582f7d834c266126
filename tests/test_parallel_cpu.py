"""World-size-2 gloo tests (CPU) of the multi-GPU partitioning logic in acfcalculator_b200/parallel.py.

The per-block arithmetic is injected from the oracle here (the product's own arithmetic is CUDA-only and is
covered on the GPU by tests/test_gpu_parity.py::test_time_blocks_add_up_on_device); what these tests pin is
the block/halo geometry, the single all-reduce, the normalisation after the sum, and the window round-robin."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, ou_series


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from acfcalculator_b200 import parallel as P
    from oracle import oracle as O
    try:
        if case == "time_blocks":
            n, m = 30011, 700
            x = ou_series(n, 40.0, seed=77)
            lo, hi, halo_end = P.time_block(n, m, world, rank)
            block = torch.from_numpy(x[lo:halo_end].copy())

            def lag_sums(b, i_count, mm):
                return torch.from_numpy(O.lag_sums_block(b.numpy(), 0, i_count, mm))
            corr = P.time_block_correlation(block, hi - lo, n, m, lag_sums_fn=lag_sums)
            np.save(os.path.join(out_dir, "corr_%d.npy" % rank), corr.numpy())
        elif case == "components":
            n, m, comps = 20000, 300, 3
            xs = [ou_series(n, 25.0, sigma=50.0, seed=20260201 + c) for c in range(comps)]
            lo, hi, halo_end = P.time_block(n, m, world, rank)
            blocks = [torch.from_numpy(x[lo:halo_end].copy()) for x in xs]

            def batch(bs, i_count, mm):
                return torch.stack([torch.from_numpy(O.lag_sums_block(b.numpy(), 0, i_count, mm)) for b in bs])
            corr = P.time_block_correlation_batch(blocks, hi - lo, n, m, batch, P._host_normalise)
            np.save(os.path.join(out_dir, "comp_%d.npy" % rank), corr.numpy())
        elif case == "windows":
            nw, m = 5, 200
            mine = P.shard_items(nw, world, rank)
            local = {}
            for w in mine:
                x = ou_series(4000 + w, 10.0 + w, sigma=0.25, mu=-32.0 + w, seed=20260101 + w)
                r = O.acf_pipeline(x, m, 2.0, 0.05)
                local[w] = (r["var"], r["acf_int"], int(r["cutoff_index"]))
            got = P.gather_results(local, nw)
            if rank == 0:
                np.save(os.path.join(out_dir, "windows.npy"), np.array(got))
            else:
                assert got is None
    finally:
        dist.destroy_process_group()


def _spawn(case, tmp_path, world=2):
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)


def test_block_geometry():
    from acfcalculator_b200 import parallel as P
    n, m = 1000003, 5000
    for world in (1, 2, 3, 8):
        edges = [P.time_block(n, m, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for r in range(world - 1):
            assert edges[r][1] == edges[r + 1][0]                    # blocks tile [0, n)
            assert edges[r][2] == min(n, edges[r][1] + m - 1)        # halo of m-1 samples
        assert edges[-1][2] == n
    assert P.shard_items(7, 3, 1) == [1, 4]
    assert sorted(sum((P.shard_items(64, 8, r) for r in range(8)), [])) == list(range(64))


def test_time_block_allreduce_matches_whole_series(oracle, tmp_path):
    _spawn("time_blocks", tmp_path)
    x = ou_series(30011, 40.0, seed=77)
    ref = oracle.calc_correlation(x, 700)
    c0, c1 = np.load(tmp_path / "corr_0.npy"), np.load(tmp_path / "corr_1.npy")
    assert np.array_equal(c0, c1)                                    # every rank holds the same result
    assert np.max(np.abs(c0 - ref)) <= 1e-12 * ref[0]


def test_component_batch_uses_one_allreduce(oracle, tmp_path):
    _spawn("components", tmp_path)
    got = np.load(tmp_path / "comp_0.npy")
    assert np.array_equal(got, np.load(tmp_path / "comp_1.npy"))
    for c in range(3):
        ref = oracle.calc_correlation(ou_series(20000, 25.0, sigma=50.0, seed=20260201 + c), 300)
        assert np.max(np.abs(got[c] - ref)) <= 1e-12 * ref[0]


def test_window_round_robin_gathers_everything(oracle, tmp_path):
    _spawn("windows", tmp_path)
    got = np.load(tmp_path / "windows.npy")
    for w in range(5):
        x = ou_series(4000 + w, 10.0 + w, sigma=0.25, mu=-32.0 + w, seed=20260101 + w)
        r = oracle.acf_pipeline(x, 200, 2.0, 0.05)
        assert got[w][0] == r["var"] and got[w][1] == r["acf_int"] and int(got[w][2]) == r["cutoff_index"]
