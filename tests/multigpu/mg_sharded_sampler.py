"""N-GPU check of the view-sharded sampler (run under torchrun): every rank denoises its own views of a V-view run
(fresh pairing per UNet call, device F recompute, partner rows over NVLink peer memory, re-noise drawn full-size and
sliced) and must reproduce the corresponding views of the un-sharded run of the same loop on one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 \
        scripts/test_sampler_2gpu.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import torch
import torch.distributed as dist

from cvd_b200 import sampler, sharding, sync_unet

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
shard = sharding.ViewShard()
peer = sharding.PeerExchange(shard)

V_loc, feat, steps, guidance = 2, 8, 3, 4.0
V = V_loc * world
ek = dict(sync_unet.EPI_KWARGS, epi_rand_slope_ff=False)        # the slope draw is per call and identical on all ranks anyway


def build():
    torch.manual_seed(7)
    return sync_unet.SyncPathUNet(levels=2, epi_kwargs=ek).randomize_zero_init(7).to(dev).eval()


g = torch.Generator().manual_seed(42)
lat = torch.randn(V, 4, 16, feat, feat, generator=g).to(dev)
text = (torch.randn(2, 77, 768, generator=g) * 0.5).repeat(V, 1, 1).to(dev)
poses = [(torch.randn(V, c, 16, feat >> i, feat >> i, generator=g) * 0.5).to(dev) for i, c in enumerate((320, 640))]
c2w, K = sampler.multiview_cameras(V, 16, "circle", 0.5, 256)
loc = slice(rank * V_loc, (rank + 1) * V_loc)


def run(net, lat_in, text_in, poses_in, sh):
    np.random.seed(123)
    torch.manual_seed(123)
    with torch.no_grad():
        return sampler.sample_multiview(net, lat_in, text_in, sampler.DDIMSchedule(device=dev), steps, guidance, aux_c2w=c2w,
                                        aux_K_mats=K, pose_embedding_features=poses_in, multistep=2, shard=sh)


ref = run(build(), lat.clone(), text, poses, None)                      # un-sharded, every rank computes all V views
net = build()
n_proc = sampler.install_sharding(net, shard, peer)
out = run(net, lat[loc].clone(), text[2 * rank * V_loc:2 * (rank + 1) * V_loc], [p[loc] for p in poses], shard)
torch.cuda.synchronize()
err = float((out - ref[loc]).norm() / ref[loc].norm())
print(f"rank {rank}/{world}: {n_proc} processors sharded; sharded vs un-sharded final latents rel-L2 = {err:.3e}", flush=True)
assert torch.isfinite(out).all() and err < 2e-2, err
dist.barrier()
dist.destroy_process_group()
if rank == 0:
    print("SHARDED SAMPLER OK")
