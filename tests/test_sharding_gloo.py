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


def _grad_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cvd_b200 import training
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.GELU(), torch.nn.Linear(40, 40), torch.nn.Linear(40, 8))
        net[1].requires_grad_(False)
        for p in net[2].parameters():              # a frozen layer in the middle, and one parameter without a gradient
            p.requires_grad_(False)
        params = training.trainable_epi_parameters(net)
        assert len(params) == 4
        g = torch.Generator().manual_seed(100 + rank)           # every rank its own batch (train_epi_control.py:283-290)
        x = torch.randn(16, 24, generator=g)
        net(x).pow(2).mean().backward()
        local = [p.grad.clone() for p in params]
        n_b = training.allreduce_gradients(params, comm_dtype=torch.float32, bucket_bytes=4096)
        assert n_b >= 2                                          # several buckets at this bucket size
        ret[rank] = ([t.numpy() for t in local], [p.grad.numpy().copy() for p in params])
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world2_equals_mean_of_rank_gradients():
    """cvd_b200.training.allreduce_gradients over gloo (fp32 wire format: exact): every rank ends with the mean of the
    ranks' local gradients, bucketed in reverse parameter order, frozen parameters untouched."""
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_grad_worker, args=(2, port, ret), nprocs=2, join=True)
    (l0, r0), (l1, r1) = ret[0], ret[1]
    for a, b, m0, m1 in zip(l0, l1, r0, r1):
        mean = (a + b) / 2
        assert np.allclose(m0, mean, atol=1e-7) and np.allclose(m1, mean, atol=1e-7)
    from cvd_b200 import training
    ps = [torch.nn.Parameter(torch.zeros(n)) for n in (10, 300, 20, 500)]
    for p in ps:
        p.grad = torch.zeros_like(p)
    sizes = [[p.numel() for p in b] for b in training.gradient_buckets(ps, 1200, 2)]
    assert sizes == [[500, 20], [300, 10]]                       # reverse parameter order, greedy 600-element buckets


class _ToyAdaptor(torch.nn.Module):
    """Stand-in with PoseAdaptor's call signature: a per-pixel linear epsilon predictor (returns (pred, auxiliary=None))."""

    def __init__(self):
        super().__init__()
        torch.manual_seed(3)
        self.mix = torch.nn.Conv3d(4, 4, 1)
        self.frozen = torch.nn.Conv3d(4, 4, 1).requires_grad_(False)

    def forward(self, noisy_latents, timesteps, encoder_hidden_states, pose_embedding, F_mats=None, H_mats=None):
        return self.mix(self.frozen(noisy_latents)) + encoder_hidden_states.mean(), None


def _step_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cvd_b200 import sampler, training
        net = _ToyAdaptor()
        opt = torch.optim.SGD(training.trainable_epi_parameters(net), lr=0.5)
        g = torch.Generator().manual_seed(200 + rank)            # every rank its own batch
        lat, noise = torch.randn(2, 4, 3, 4, 4, generator=g), torch.randn(2, 4, 3, 4, 4, generator=g)
        ts = torch.tensor([900 - 100 * rank, 17 + rank])
        loss, total = training.training_step(net, opt, sampler.DDIMSchedule(), lat, noise, ts, torch.ones(2, 3, 5),
                                             max_grad_norm=0.05, comm_dtype=torch.float32)
        ret[rank] = (float(loss), float(total), [p.detach().numpy().copy() for p in net.parameters()])
    finally:
        dist.destroy_process_group()


def test_training_step_world2_equals_one_process_on_the_joint_batch():
    """cvd_b200.training.training_step under gloo, world 2: gradients are averaged across the ranks BEFORE clipping and the
    optimizer step (what DistributedDataParallel does for train_epi_control.py:349,612-630), so both ranks end with the
    same parameters -- those one process reaches with the mean of the two ranks' losses."""
    from cvd_b200 import sampler, training
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_step_worker, args=(2, port, ret), nprocs=2, join=True)
    (loss0, tot0, p0), (loss1, tot1, p1) = ret[0], ret[1]
    assert abs(tot0 - tot1) < 1e-6 and all(np.array_equal(a, b) for a, b in zip(p0, p1)) and loss0 != loss1
    net = _ToyAdaptor()
    params = training.trainable_epi_parameters(net)
    sch = sampler.DDIMSchedule()
    total_loss = 0.0
    for rank in range(2):
        g = torch.Generator().manual_seed(200 + rank)
        lat, noise = torch.randn(2, 4, 3, 4, 4, generator=g), torch.randn(2, 4, 3, 4, 4, generator=g)
        ts = torch.tensor([900 - 100 * rank, 17 + rank])
        pred, _ = net(sch.add_noise(lat, noise, ts), ts, torch.ones(2, 3, 5), None)
        total_loss = total_loss + torch.nn.functional.mse_loss(pred, noise) / 2
    total_loss.backward()
    tot = torch.nn.utils.clip_grad_norm_(params, 0.05)
    torch.optim.SGD(params, lr=0.5).step()
    assert abs(float(tot) - tot0) < 1e-5 and float(tot) > 0.05          # clipping active
    for a, b in zip(p0, net.parameters()):
        assert np.allclose(a, b.detach().numpy(), atol=1e-6)
