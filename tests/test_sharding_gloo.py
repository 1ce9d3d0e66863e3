"""CPU, world_size 2 over gloo: host-side logic of the view-sharded epi attention
(cvd_b200/sharding.py): row ownership, global partner indices, the K/V^T all-gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import golden_inputs as gi


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cvd_b200
        from cvd_b200.sharding import ViewShard
        shard = ViewShard()
        R_glob, N, C = 64, 4, 8                      # 4 views x 16 frames
        rows = shard.local_rows(R_glob)
        assert (rows.start, rows.stop) == (rank * 32, rank * 32 + 32)
        g = torch.Generator().manual_seed(0)
        k_glob = torch.randn(R_glob, N, C, generator=g)
        k_all = shard.gather_rows(k_glob[rows])       # what every rank sees after the exchange
        assert torch.equal(k_all, k_glob)
        # view pairing v <-> 3-v as the V-view sampler would pass it (global rows)
        kv = gi.view_pair_index(4)
        local = shard.local_kv_index(kv, R_glob)
        assert local.tolist() == kv[rows].tolist()
        # the rows this rank attends to are exactly the un-sharded gather
        assert torch.equal(k_all[torch.from_numpy(local)], k_glob[torch.from_numpy(kv)][rows])
        # default partner (swap halves) in global numbering: rank 0 reads rank 1's rows and vice versa
        sw = shard.swap_halves_index(R_glob)
        assert sw.tolist() == list(range(32, 64)) if rank == 0 else sw.tolist() == list(range(0, 32))
        # two partners per row: [2*R_glob] global index -> [2*R_local]
        multi = np.concatenate([kv, np.arange(R_glob)])
        lm = shard.local_kv_index(multi, R_glob)
        assert lm.shape[0] == 64 and lm[32:].tolist() == list(range(rank * 32, rank * 32 + 32))
        # the processor resolves indices against the gathered row count
        proc = cvd_b200.EpiAttnProcessor(hidden_size=C)
        proc.shard = shard
        idx = proc.partner_index(32, local, torch.device("cpu"))
        assert idx.dtype == torch.int32 and int(idx.max()) < R_glob and idx.tolist() == local.tolist()
        assert proc.partner_index(32, None, torch.device("cpu")).tolist() == sw.tolist()
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_view_shard_world2():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
