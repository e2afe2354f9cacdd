"""CPU, world_size 2 over gloo: member sharding and the ensemble-statistics all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from watergap2_b200.ensemble import ensemble_mean_var, routing_levels, shard_by_basin, shard_members, subgrid_inputs, tile_inputs


def test_shard_members_partition():
    for total in (0, 1, 7, 256, 1024):
        for world in (1, 2, 3, 8):
            blocks = [shard_members(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        shard_members(8, 2, 2)


def test_shard_by_basin_keeps_basins_whole(world3000, oracle_lib):
    w = world3000
    t = oracle_lib.rout_prepare(w.flowdir, w.row, w.col, w.gcrc.T)
    rank = shard_by_basin(t["basins2"], 4)
    b = t["basins2"]
    for bid in np.unique(b[b > 0]):
        assert len(set(rank[b == bid])) == 1
    down = t["outflow_cell"]
    has = down > 0
    assert (rank[has] == rank[down[has] - 1]).all()  # water never crosses a rank boundary
    load = np.bincount(rank, minlength=4)
    assert load.max() <= 1.3 * load.mean() + np.bincount(b[b > 0]).max()


def test_subgrid_inputs_keep_order_and_levels(world3000, oracle_lib):
    """basin shards as stand-alone grids: ranks stay level-major, relative order and levels unchanged"""
    from oracle import wg_init
    w = world3000
    ini = wg_init.derive(w)
    topo = ini["_topology"]
    rank = shard_by_basin(oracle_lib.rout_prepare(w.flowdir, w.row, w.col, w.gcrc.T)["basins2"], 2)
    ro, dc = np.asarray(topo["rout_order"]), np.asarray(topo["outflow_cell"])
    def levels(ro, dc):
        lvl = np.zeros(ro.size, int)
        for n in np.argsort(ro):
            if dc[n] > 0:
                lvl[dc[n] - 1] = max(lvl[dc[n] - 1], lvl[n] + 1)
        return lvl
    full = levels(ro, dc)
    seen = 0
    for r in range(2):
        cells = np.nonzero(rank == r)[0]
        f, sro, sdc = subgrid_inputs(ini, ro, dc, cells)
        assert sorted(sro.tolist()) == list(range(1, cells.size + 1))
        assert (np.argsort(sro) == np.argsort(ro[cells])).all()
        assert np.array_equal(levels(sro, sdc), full[cells])
        assert f["area"].shape[0] == cells.size and f["params"].shape == (26, cells.size)
        assert np.asarray(f["snow_bands"]).size == cells.size * 101 and np.asarray(f["lct_ddf"]).size == 18
        seen += cells.size
    assert seen == w.ng
    with pytest.raises(ValueError):
        subgrid_inputs(ini, ro, dc, np.nonzero(dc > 0)[0][:5])  # cells cut out of their basins


def test_tile_inputs_is_a_valid_level_major_grid(world1000):
    from oracle import wg_init
    ini = wg_init.derive(world1000)
    topo = ini["_topology"]
    ro, dc = np.asarray(topo["rout_order"]), np.asarray(topo["outflow_cell"])
    f, tro, tdc = tile_inputs(ini, ro, dc, 3)
    ng = ro.size
    assert sorted(tro.tolist()) == list(range(1, 3 * ng + 1))
    lvl = routing_levels(tro, tdc)
    assert np.array_equal(lvl, np.tile(routing_levels(ro, dc), 3))
    assert (np.diff(lvl[np.argsort(tro)]) >= 0).all()  # level-major, as rout_order's sweeps number the cells
    has = tdc > 0
    assert ((tdc[has] - 1) // ng == np.nonzero(has)[0] // ng).all()  # no water between the copies
    assert f["params"].shape == (26, 3 * ng) and np.asarray(f["snow_bands"]).size == 3 * ng * 101
    assert np.array_equal(np.asarray(f["area"])[ng:2 * ng], np.asarray(ini["area"]))


def _worker(rank, world, port, total, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.from_numpy(np.random.default_rng(7).standard_normal((total, n)))
    first, count = shard_members(total, world, rank)
    mean, var = ensemble_mean_var(full[first:first + count].clone(), total)
    if rank == 0:
        torch.save({"mean": mean, "var": var, "ref_mean": full.mean(0), "ref_var": full.var(0, unbiased=False)}, out)
    dist.destroy_process_group()


def test_ensemble_statistics_world_size_2(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "stats.pt")
    mp.spawn(_worker, args=(2, port, 13, 257, out), nprocs=2, join=True)
    r = torch.load(out)
    assert torch.allclose(r["mean"], r["ref_mean"], rtol=0, atol=1e-14)
    assert torch.allclose(r["var"], r["ref_var"], rtol=1e-12, atol=1e-14)


def _crit_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from watergap2_b200 import calibration as cal
    from watergap2_b200.ensemble import shard_members
    nsets, nyears = 7, 3
    lo, cnt = shard_members(nsets, world, rank)
    hi = lo + cnt
    rng = np.random.default_rng(5)
    measured = rng.uniform(5., 9., nyears).astype(np.float32)
    sims = rng.uniform(4., 10., (nsets, nyears)).astype(np.float32)  # every rank draws the same table, keeps its shard
    mine = [[cal.criteria(measured, sims[k])] for k in range(lo, hi)]
    allc = cal.gather_criteria(mine)
    if rank == 0:
        ref = [[cal.criteria(measured, sims[k])] for k in range(nsets)]
        torch.save({"ok": allc == ref, "n": len(allc), "best": cal.best_member(allc, 0),
                    "best_ref": int(np.argmin([r[0]["rel_difference"] for r in ref]))}, out)
    dist.destroy_process_group()


def test_sweep_criteria_gather_world_size_2(tmp_path):
    """the calibration sweep over ranks: every rank evaluates the criteria of its own parameter sets, gather_criteria returns the
    list of all sets in set order (config 3 of BASELINE.json: 1024 sets sharded over the GPUs)"""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "crit.pt")
    mp.spawn(_crit_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ok"] and r["n"] == 7 and r["best"] == r["best_ref"]
