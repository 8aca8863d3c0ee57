"""N > 1 host logic on CPU: two gloo ranks shard one plan and sum their count tables.
(The GPU kernels cannot run here; the per-rank tables are stand-ins derived from the
rank's own tiles, which is exactly what the reduce step has to add up.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import make_params, small_spec


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from process_b200 import _lib as L
        from process_b200 import api
        from process_b200.synth import synth_forest
        f = synth_forest(small_spec(3))
        fl = L.Flat(f)
        assert api._shard() == (rank, world)
        info_all, t_all = fl.plan(make_params(coverage=30.0, purity=0.7))
        info, t = fl.plan(make_params(coverage=30.0, purity=0.7, shard_rank=rank, shard_count=world))
        assert info.n_templates_total == info_all.n_templates_total
        # stand-in tables: templates of my tiles per (sample, chromosome)
        S = info.n_out_samples
        occ = np.zeros((S, f.n_chr), np.uint32)
        np.add.at(occ, (t["sample"], t["chr"]), t["templates"])
        cov = occ * 2
        occ_sum, cov_sum = api._reduce_over_ranks(occ, cov)
        want = np.zeros((S, f.n_chr), np.uint32)
        np.add.at(want, (t_all["sample"], t_all["chr"]), t_all["templates"])
        assert np.array_equal(occ_sum, want) and np.array_equal(cov_sum, want * 2)
        # every tile is owned by exactly one rank
        ids = torch.zeros(int(info_all.n_tiles_total), dtype=torch.int32)
        ids[torch.from_numpy(t["id"].astype(np.int64))] = 1
        dist.all_reduce(ids)
        owned = torch.zeros_like(ids)
        owned[torch.from_numpy(t_all["id"].astype(np.int64))] = 1
        assert torch.equal(ids, owned)
        # the shards are balanced (on the tiles' cost: templates x a locus-density term), so the loads are close
        load = torch.tensor([float(info.n_templates)])
        lo, hi = load.clone(), load.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert hi.item() / lo.item() < 1.15
        # seed=None: every rank draws from its OWN RNG state; the call must still plan one job.  The seed is resolved
        # on rank 0 and broadcast, and the shards planned with it partition the tile grid
        np.random.seed(1000 + rank)
        own = api._resolve_seed(None)
        seeds = [None] * world
        dist.all_gather_object(seeds, own)
        assert len(set(seeds)) == world  # the ranks did draw different seeds
        agreed = api._agree_over_ranks(own, None, list(f.sample_names))
        assert agreed == seeds[0]
        info_all, t_all = fl.plan(make_params(coverage=30.0, purity=0.7, seed=agreed))
        info, t = fl.plan(make_params(coverage=30.0, purity=0.7, seed=agreed, shard_rank=rank, shard_count=world))
        ids = torch.zeros(int(info_all.n_tiles_total), dtype=torch.int32)
        ids[torch.from_numpy(t["id"].astype(np.int64))] = 1
        dist.all_reduce(ids)
        owned = torch.zeros_like(ids)
        owned[torch.from_numpy(t_all["id"].astype(np.int64))] = 1
        assert torch.equal(ids, owned)
        # ranks that disagree on the sample partition are refused
        try:
            api._agree_over_ranks(own, np.full(f.n_leaves, rank, np.uint32), ["a", "b"])
            raise AssertionError("differing partitions were accepted")
        except ValueError:
            pass
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_gloo_ranks_shard_and_reduce(tmp_path, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
