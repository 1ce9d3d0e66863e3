#!/usr/bin/env python
"""Benchmark of the CVD epipolar cross-view attention hot path (BASELINE.json metric:
epi-attn TFLOP/s and % of bf16 tensor peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE config 2, two-video inference on one B200): R = 64 rows (2 views x 2 CFG
branches x 16 frames); one STEP = the 40 epi-attention calls of one UNet forward, 10 each at
(N, C) = (1024, 320), (256, 640), (64, 1280), (16, 1280), 8 heads, real pose-pair geometry,
random-init weights, synthetic unit-variance activations.  Algorithmic work per step =
sum 4*R*N*N*C = 15.32 GFLOP * R (softmax, bias and MMA padding excluded).

  value : TFLOP/s of the step with q / k / v^T / F_mats resident in HBM (row-table prepass + the
          fused tcgen05 kernel per call; distinct buffers per call, 1.9 GB per step >> L2).
  e2e   : the same metric through the drop-in layer API (EpiSelfAttention.forward) with HOST
          buffers: per call, H2D of the fp32 hidden states from pinned memory, projections, bias
          prepass, fused attention, output projection, D2H of the result.
  --impl reference : the reference's CPU path (oracle port of EpiSelfAttention.forward, fp32,
          all host threads) on a bounded sample of the same step (one call per shape).
At N > 1 the views shard across ranks (weak scaling: 64 rows per GPU, 2N views); K and V^T live in
symmetric memory and every call reads exactly its partner rows across NVLink after one device-side
barrier: by the copy engines in row chunks, the smallest first, each chunk's attention running under
the transfer of the next (--exchange dma, default; latents <= 8x8 are read inside the attention
kernel by unicast TMA over NVLink), by our pull kernel (--exchange peer), or as an NCCL all-gather of
every rank's rows (--exchange allgather, the baseline).  At N > 1 the step is timed twice: issued eagerly
(two-chunk plan, phase A) and -- last, with every other number of the line already safe -- replayed as ONE CUDA
graph (wave-sized chunks, phase B); the replay becomes the line's value only if every rank captured it, its
outputs equal the eager step's bit for bit and it is faster (config.exchange_issue says which).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

SHAPES = [(32, 320), (16, 640), (8, 1280), (4, 1280)]   # (feat, C) of epi_module_resolutions [1,2,4,8] at 256x256
CALLS_PER_SHAPE = 10                                     # 5 epi modules x 2 attention blocks per resolution
HEADS = 8
R_PER_GPU = 64                                           # 2 views x 2 cfg x 16 frames
F_MAT_SIZE = 256
DENOISE_FEAT, DENOISE_LEVELS = 32, 4                     # latent size / UNet levels of the denoise measurement (256x256 images)
EXCHANGE = "NVLink copy engines in row chunks (smallest first) under the attention, K and V^T on two streams; grids <= 8x8 read the peer's K/V^T in-kernel (unicast TMA over NVLink); issue: config.exchange_issue"


def step_flops(R, shapes=SHAPES, calls=CALLS_PER_SHAPE):
    return sum(calls * 4.0 * R * (f * f) ** 2 * C for f, C in shapes)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["bf16_tflops"]), float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst)"
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the
    committed ncu --set full capture of this round (profiles/), or None."""
    path = os.path.join(ROOT, "profiles", "r01_epi_attn_d40_ncu_full_v7.json")
    try:
        d = json.load(open(path))
        rd = float(d["dram__bytes_read.sum"].split()[0]) * 1e6
        wr = float(d["dram__bytes_write.sum"].split()[0]) * 1e6
        return rd + wr
    except Exception:
        return None


def golden_F():
    g = np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))
    return torch.from_numpy(g["F_mats"])          # [32,3,3]: rows 0-15 view A->B, 16-31 view B->A


def cfg_F_mats(n_views_pairs=1):
    """[R,3,3] in the sampler's CFG layout [U-src, C-src, U-tgt, C-tgt] (pipeline_animation_epi.py:565-571)."""
    Fm = golden_F()
    src, tgt = Fm[:16], Fm[16:]
    return torch.cat([src, src, tgt, tgt], dim=0).repeat(n_views_pairs, 1, 1).contiguous()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: oracle port of the reference's CPU path
# ------------------------------------------------------------------------------------------
def cpu_reference_step(R, reps, warm):
    """Times the oracle's EpiSelfAttention.forward (fp32, all host threads) on a bounded sample of
    the step: ONE call per shape.  Returns (seconds per sample step, flops per sample step, cores)."""
    import golden_inputs as gi
    from oracle import epi_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Fm = cfg_F_mats(R // 64)
    cases = []
    for feat, C in SHAPES:
        sd = gi.attn_state_dict(C, 10 + feat)
        x = torch.randn(R, feat * feat, C, generator=torch.Generator().manual_seed(feat))
        cases.append((x, sd, Fm))
    flops = step_flops(R, calls=1)

    def one():
        with torch.no_grad():
            for x, sd, F_ in cases:
                O.epi_self_attention(x, sd, "", HEADS, F_, F_MAT_SIZE, theta=0.5)

    for _ in range(warm):
        one()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    return times, flops, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    R = R_PER_GPU * args.gpus
    warm = args.warmup
    times, flops, cores = cpu_reference_step(R, args.steps, warm)
    total = sum(times)
    val = flops * len(times) / total / 1e12
    sample = f"1 of the {CALLS_PER_SHAPE} calls per shape per step (4 calls, {flops / 1e9:.1f} GFLOP), fp32, {cores} threads"
    line = {
        "impl": "reference", "metric": "epi_attn_tflops", "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    return {"workload": "CVD two-video inference (BASELINE config 2): 40 epi-attention calls of one UNet forward, "
                        "R=64 rows per GPU (2 views x 2 cfg x 16 frames), (N,C) = (1024,320),(256,640),(64,1280),(16,1280) x10, 8 heads",
            "rows_per_gpu": R_PER_GPU, "views_total": 2 * n_gpus, "flops_per_step": step_flops(R_PER_GPU * n_gpus),
            "parallelism": "1 GPU" if n_gpus == 1 else f"views sharded over {n_gpus} GPUs, partner K/V^T rows pulled per call ({EXCHANGE})",
            "l2_policy": "inputs larger than L2: distinct q/k/v^T buffers per call (1.9 GB per step)"}


def denoise_bench(dev, world, rank, shard, peer, steps):
    """The second half of the BASELINE metric on the path this repository owns: denoising steps/s of the sampler
    loops (cvd_b200/sampler.py) on the hot-path skeleton UNet (cvd_b200/sync_unet.py: the 20 temporal + 20 epi
    modules of the reference UNet at 256x256; the SD-1.5 spatial blocks are out of scope and NOT in this number).
    One step = one CFG-batched UNet forward (R = 64 rows per GPU) + the DDIM update.  N = 1: the two-video loop on the
    real pose pair; N > 1: the V = 2N-view loop (fresh pairing + device F recompute per call, partner rows over NVLink)."""
    from cvd_b200 import sampler, sync_unet
    torch.manual_seed(0)
    np.random.seed(0)
    net = sync_unet.SyncPathUNet(levels=DENOISE_LEVELS).randomize_zero_init(1).to(dev).eval()
    chans, ft = sync_unet.CHANNELS[:DENOISE_LEVELS], DENOISE_FEAT
    V_loc, V = 2, 2 * world
    g = torch.Generator(device=dev).manual_seed(42)
    lat = torch.randn(V_loc, 4, 16, ft, ft, device=dev, generator=g)
    poses = [torch.randn(V_loc, c, 16, ft >> i, ft >> i, device=dev, generator=g) for i, c in enumerate(chans)]
    sch = sampler.DDIMSchedule(device=dev)
    if world == 1:
        text = torch.randn(2, 77, 768, device=dev, generator=g)
        Fm = golden_F().reshape(2, 16, 3, 3).to(dev)

        # the whole UNet forward replays as ONE CUDA graph (about 1500 launches per forward otherwise); the timestep
        # enters as a device scalar so that every step reuses the same graph
        import cvd_b200
        gnet = cvd_b200.GraphedModule(net)

        def unet(x, t, **kw):
            return gnet(x, torch.full((), float(t), device=dev), **kw)

        def run(n, graphed=True):
            return sampler.sample_pair(unet if graphed else net, lat, Fm, text, sch, n, 8.5, pose_embedding_features=poses)
    else:
        sampler.install_sharding(net, shard, peer)
        text = torch.randn(2, 77, 768, device=dev, generator=g).repeat(V_loc, 1, 1)
        c2w, K = sampler.multiview_cameras(V, 16, "circle", 0.5, 256)

        def run(n):
            return sampler.sample_multiview(net, lat, text, sch, n, 8.5, aux_c2w=c2w, aux_K_mats=K,
                                            pose_embedding_features=poses, shard=shard)
    with torch.no_grad():
        run(3)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = run(steps)
        e1.record()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        ms_eager = None
        if world == 1:      # the same loop without the graph, for the record
            run(2, graphed=False)
            torch.cuda.synchronize()
            e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e2.record()
            run(steps, graphed=False)
            e3.record()
            torch.cuda.synchronize()
            ms_eager = e2.elapsed_time(e3) / steps
    assert torch.isfinite(out).all()
    by_views = config4 = None
    if world > 1 and 24 % world == 0:
        def config4():
            """BASELINE config 4: 6-view 'upper_hemi' generation, accumulate_step 2 -> 6 x 2 (cfg) x 2 (acc) = 24 independent
            UNet evaluations per update, sharded as UNITS (24 / world per GPU: 3 on 8 GPUs), partner rows through the layers'
            NVLink exchange, one small all-reduce of the weighted predictions per update (sampler.sample_multiview_units).
            Returned to the caller, which runs it LAST (its code had not run on hardware when this was written).  A Python
            exception is reported in the line (every rank runs the same host logic, so it fails on all ranks or none)."""
            try:
                V6 = 6
                g6 = torch.Generator(device=dev).manual_seed(43)
                lat6 = torch.randn(V6, 4, 16, ft, ft, device=dev, generator=g6)
                poses6 = [torch.randn(V6, c, 16, ft >> i, ft >> i, device=dev, generator=g6) for i, c in enumerate(chans)]
                text6 = torch.randn(2, 77, 768, device=dev, generator=g6).repeat(V6, 1, 1)
                c2w6, K6 = sampler.multiview_cameras(V6, 16, "upper_hemi", 0.5, 256)

                def run6(n):
                    return sampler.sample_multiview_units(net, lat6, text6, sch, n, 8.5, aux_c2w=c2w6, aux_K_mats=K6,
                                                          pose_embedding_features=poses6, accumulate_step=2, shard=shard)
                with torch.no_grad():
                    run6(2)
                    torch.distributed.barrier()
                    torch.cuda.synchronize()
                    a6, b6 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a6.record()
                    o6 = run6(steps)
                    b6.record()
                    torch.distributed.barrier()
                    torch.cuda.synchronize()
                ms6 = torch.tensor([a6.elapsed_time(b6) / steps], device=dev, dtype=torch.float64)
                torch.distributed.all_reduce(ms6, op=torch.distributed.ReduceOp.MAX)
                return {"config4": {"steps_per_s": 1e3 / float(ms6.item()), "views": V6, "accumulate_step": 2,
                                    "units_per_gpu": 24 // world, "rig": "upper_hemi", "finite": bool(torch.isfinite(o6).all()),
                                    "loop": "sampler.sample_multiview_units ((view, cfg, acc) units sharded, one all-reduce per update)"}}
            except RuntimeError as exc:
                if "CUDA" in str(exc):      # not survivable here: the caller prints the held line and leaves
                    raise
                return {"config4": {"error": f"{type(exc).__name__}: {exc}"}}
            except Exception as exc:
                return {"config4": {"error": f"{type(exc).__name__}: {exc}"}}
    if world == 1:
        # BASELINE metric "denoise steps/s at 2/4/6 views": the V-view loop (fresh pairing, device F recompute and kv_index
        # gather per UNet call) with all V views on this one GPU (R = 32 V rows), multistep = 1
        by_views = {"2": {"steps_per_s": 1e3 / (e0.elapsed_time(e1) / steps), "rows": 64, "launch": "one CUDA graph per UNet forward"}}
        for Vn, rig in ((4, "circle"), (6, "upper_hemi")):
            lat_v = torch.randn(Vn, 4, 16, ft, ft, device=dev, generator=g)
            poses_v = [torch.randn(Vn, c, 16, ft >> i, ft >> i, device=dev, generator=g) for i, c in enumerate(chans)]
            text_v = torch.randn(2, 77, 768, device=dev, generator=g).repeat(Vn, 1, 1)
            c2w, K = sampler.multiview_cameras(Vn, 16, rig, 0.5, 256)

            def run_v(n):     # kv_index as a device tensor: ONE graph serves every pairing
                return sampler.sample_multiview(unet, lat_v, text_v, sch, n, 8.5, aux_c2w=c2w, aux_K_mats=K,
                                                pose_embedding_features=poses_v, kv_on_device=True)
            with torch.no_grad():
                run_v(2)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                o = run_v(steps)
                b.record()
                torch.cuda.synchronize()
            assert torch.isfinite(o).all()
            by_views[str(Vn)] = {"steps_per_s": 1e3 / (a.elapsed_time(b) / steps), "rows": 32 * Vn,
                                 "launch": "one CUDA graph per UNet forward (kv_index on the device)", "rig": rig}
    return e0.elapsed_time(e1) / steps, V, ms_eager, by_views, config4


# SD-1.5 + AnimateDiff v3 + CameraCtrl + CVD: the UNet of configs/inference_config.yaml (unet_additional_kwargs) on top of the
# SD-1.5 config.json (block_out_channels, cross_attention_dim 768, 8 heads, 2 layers per block)
FULL_UNET_KWARGS = dict(
    sample_size=32, in_channels=4, out_channels=4, block_out_channels=(320, 640, 1280, 1280), layers_per_block=2,
    cross_attention_dim=768, attention_head_dim=8, norm_num_groups=32,
    use_motion_module=True, motion_module_resolutions=(1, 2, 4, 8), motion_module_mid_block=False, motion_module_type="Vanilla",
    motion_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Temporal_Self", "Temporal_Self"],
                              temporal_position_encoding=True, temporal_position_encoding_max_len=32,
                              temporal_attention_dim_div=1, zero_initialize=False),
    use_epi_module=True, epi_module_resolutions=(1, 2, 4, 8), epi_module_mid_block=False,
    epi_module_kwargs=dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Epi_Self", "Epi_Self"],
                           epi_position_encoding=True, epi_position_encoding_feat_max_size=64,
                           epi_position_encoding_F_mat_size=256, epi_rand_slope_ff=True, zero_initialize=True))
FULL_UNET_TEXT_LEN = 77


def full_unet_bench(dev, steps):
    """BASELINE metric (ii) on the WHOLE UNet: denoising steps/s of the two-video loop
    (sampler.sample_pair: CFG batch of 4 videos = 64 rows at 32x32 latents, 256x256 images) on
    cvd_b200.unet.UNet3DConditionModelPoseCond -- SD-1.5 spatial blocks (library convolutions / GEMMs under bf16 autocast)
    with this package's motion and epi modules, random-init weights (zero-initialised pieces re-initialised so that the sync
    path is visible), one CUDA graph per UNet forward -- and, beside it, ONE forward of the reference's own PyTorch GPU path
    on the same parameter tree: the same UNet class in eager fp32 with oracle-backed motion / epi modules
    (oracle/module_twins.py: dense bias per call, mask x 8 heads, plain softmax attention)."""
    import cvd_b200
    from cvd_b200 import sampler
    from cvd_b200.unet import UNet3DConditionModelPoseCond
    from oracle import module_twins
    ch = FULL_UNET_KWARGS["block_out_channels"]
    ft = FULL_UNET_KWARGS["sample_size"]
    proc = dict(add_spatial=False, spatial_attn_names="attn1", add_temporal=True, temporal_attn_names="0",
                pose_feature_dimensions=list(ch), query_condition=True, key_value_condition=True, scale=1.0,
                add_spatial_lora=False, add_motion_lora=False, lora_kwargs={"lora_rank": 2, "lora_scale": 1.0},
                motion_lora_kwargs={"lora_rank": -1, "lora_scale": 1.0},
                sync_lora_kwargs={"sync_lora_rank": 0, "sync_lora_scale": 0}, spatial_extended_attention=False)
    torch.manual_seed(0)
    with torch.device(dev):                          # parameters are created (and randomly initialised) on the device
        net = UNet3DConditionModelPoseCond(**FULL_UNET_KWARGS)
        net.set_all_attn_processor(**proc)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p in net.named_parameters():      # zero_initialize / qkv_merge: visible in a random-init run
            if (name.endswith("proj_out.weight") and "epi_modules" in name) or name.endswith("qkv_merge.weight"):
                p.copy_(torch.randn(p.shape, generator=g) * (0.5 / p.shape[1] ** 0.5))
    net = net.to(dev).eval()
    n_params = sum(p.numel() for p in net.parameters())
    gd = torch.Generator(device=dev).manual_seed(42)
    lat = torch.randn(2, 4, 16, ft, ft, device=dev, generator=gd)
    text = torch.randn(2, FULL_UNET_TEXT_LEN, FULL_UNET_KWARGS["cross_attention_dim"], device=dev, generator=gd)
    poses = [torch.randn(2, c, 16, ft >> i, ft >> i, device=dev, generator=gd) * 0.5 for i, c in enumerate(ch)]
    Fm = golden_F().reshape(2, 16, 3, 3).to(dev)
    sch = sampler.DDIMSchedule(device=dev)
    gnet = cvd_b200.GraphedModule(net)

    def unet(x, t, **kw):
        return gnet(x, torch.full((), float(t), device=dev), **kw)

    def run(n, fn):
        return sampler.sample_pair(fn, lat, Fm, text, sch, n, 8.5, pose_embedding_features=poses)
    out = {"unet": "cvd_b200.unet.UNet3DConditionModelPoseCond, SD-1.5 + AnimateDiff v3 + CameraCtrl + CVD layout, random init",
           "parameters_M": n_params / 1e6, "views": 2, "rows": 64, "sd_blocks": "bf16 autocast (cuDNN / cuBLAS / SDPA)"}
    with torch.no_grad():
        run(2, unet)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = run(steps, unet)
        b.record()
        torch.cuda.synchronize()
        assert torch.isfinite(res).all()
        out["ms_per_step"] = a.elapsed_time(b) / steps
        out["steps_per_s"] = 1e3 / out["ms_per_step"]
        out["launch"] = "one CUDA graph per UNet forward"
        # the reference's own PyTorch GPU path on the same weights: eager fp32, oracle-backed sync / temporal modules
        with torch.device(dev):
            twin = UNet3DConditionModelPoseCond(**dict(FULL_UNET_KWARGS, amp=False), **module_twins.unet_factories())
            twin.set_all_attn_processor(**proc)
        twin.load_state_dict(net.state_dict(), strict=False)
        twin = twin.to(dev).eval()
        for blk in twin.modules():                   # the twin's epi modules use the horizontal first-frame rule
            if hasattr(blk, "rand_slope_on_first_frame"):
                blk.rand_slope_on_first_frame = False
        x4 = torch.cat([lat[:1], lat[:1], lat[1:], lat[1:]])
        F4 = torch.cat([Fm[:1], Fm[:1], Fm[1:], Fm[1:]])
        t4 = text.repeat(2, 1, 1)
        p4 = [torch.cat([x[:1], x[:1], x[1:], x[1:]]) for x in poses]
        with torch.device(dev):
            y_ref = twin(x4, 481, encoder_hidden_states=t4, F_mats=F4, pose_embedding_features=p4).sample
            torch.cuda.synchronize()
            a.record()
            for _ in range(2):
                y_ref = twin(x4, 481, encoder_hidden_states=t4, F_mats=F4, pose_embedding_features=p4).sample
            b.record()
            torch.cuda.synchronize()
        ms_ref = a.elapsed_time(b) / 2
        for m_ in net.modules():                     # same first-frame rule on both sides for the difference below
            if hasattr(m_, "rand_slope_on_first_frame"):
                m_.rand_slope_on_first_frame = False
        y_our = net(x4, 481, encoder_hidden_states=t4, F_mats=F4, pose_embedding_features=p4).sample
        a.record()
        for _ in range(3):
            y_our = net(x4, 481, encoder_hidden_states=t4, F_mats=F4, pose_embedding_features=p4).sample
        b.record()
        torch.cuda.synchronize()
        ms_eager = a.elapsed_time(b) / 3
        out["forward"] = {"reference_gpu_fp32_eager_ms": ms_ref, "ours_eager_ms": ms_eager,
                          "ours_graphed_ms": out["ms_per_step"], "speedup_vs_reference_gpu": ms_ref / out["ms_per_step"],
                          "rel_l2": float((y_our.double() - y_ref.double()).norm() / y_ref.double().norm())}
    return out


def glue_extra(dev, peak_bw):
    """roofline_glue : achieved HBM GB/s of every glue kernel (GroupNorm + layout, LayerNorm, residual, GEGLU, temporal
    attention, ...) at the config-2 shapes against the measured copy bandwidth (scripts/profile_glue.py).  A failure is
    reported in the line, never fatal."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import profile_glue
        rows = profile_glue.measure(dev, peak_bw)
        return [{"kernel": r["kernel"], "shape": r["shape"], "us": round(r["us"], 2), "gbs": round(r["gbs"], 1),
                 "frac": round(r["frac"], 4)} for r in rows]
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"}


def reference_gpu_extra(dev, F_rows):
    """reference_gpu : the reference's own PyTorch GPU path (oracle/gpu_reference.py: fp32 eager -- dense bias per call, NaN
    check with its host sync, mask x 8 heads, SDPA -- and the same under bf16 autocast) for ONE EpiSelfAttention.forward
    per shape on device tensors, beside the product layer on the same inputs.  Guarded like glue_extra."""
    try:
        import cvd_b200
        import golden_inputs as gi
        from oracle.gpu_reference import reference_gpu_layer

        def timed(fn, iters, warm):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                out = fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / iters, out

        per_shape, tot = [], {"fp32": 0.0, "bf16_autocast": 0.0, "ours": 0.0}
        with torch.no_grad():
            for feat, C in SHAPES:
                sd = gi.attn_state_dict(C, 10 + feat)
                sd_dev = {k: v.to(dev) for k, v in sd.items()}
                x = torch.randn(R_PER_GPU, feat * feat, C, device=dev, generator=torch.Generator(device=dev).manual_seed(feat))
                attn = cvd_b200.EpiSelfAttention(attention_mode="Epi_Self", cross_attention_dim=None, query_dim=C, heads=HEADS,
                                                 dim_head=C // HEADS, epi_position_encoding=True,
                                                 epi_position_encoding_feat_max_size=64,
                                                 epi_position_encoding_F_mat_size=F_MAT_SIZE, epi_rand_slope_ff=False)
                attn.load_state_dict(sd, strict=False)
                attn = attn.to(dev).eval()
                t32, y32 = timed(lambda: reference_gpu_layer(x, sd_dev, HEADS, F_rows, F_MAT_SIZE), 3, 1)
                t16, y16 = timed(lambda: reference_gpu_layer(x, sd_dev, HEADS, F_rows, F_MAT_SIZE, autocast=True), 3, 1)
                tour, your = timed(lambda: attn(x, F_mats=F_rows)[0], 10, 3)
                err = float((your.double() - y32.double()).norm() / y32.double().norm())
                per_shape.append({"N": feat * feat, "C": C, "reference_fp32_ms": t32, "reference_bf16_autocast_ms": t16,
                                  "ours_ms": tour, "rel_l2_vs_fp32": err})
                tot["fp32"] += t32
                tot["bf16_autocast"] += t16
                tot["ours"] += tour
                del attn, sd_dev, x, y32, y16, your
                torch.cuda.empty_cache()
        ref_gpu = {"what": "EpiSelfAttention.forward, R=64, device tensors, one call per shape (the reference's PyTorch GPU "
                           "path = oracle/gpu_reference.py: dense bias + mask x 8 heads + SDPA)",
                   "per_shape": per_shape, "sum_ms": tot,
                   "speedup_vs_fp32": tot["fp32"] / tot["ours"], "speedup_vs_bf16_autocast": tot["bf16_autocast"] / tot["ours"]}
    except Exception as exc:
        ref_gpu = {"error": f"{type(exc).__name__}: {exc}"}
    return ref_gpu


EXTRA_KEYS = ("roofline_glue", "reference_gpu", "full_unet")
EXTRAS_TIMEOUT_S = 300
LATE_TIMEOUT_S = 300         # N > 1: graph replay + config-4 measurement together (seconds of GPU work when healthy)


def hard_exit(code):
    """Leave without running any CUDA / process-group teardown (used after a CUDA error or a hang in a late measurement)."""
    os._exit(code)


def extras_measurements(dev, steps, denoise):
    """The explanatory measurements of the N = 1 line, one (key, result) at a time: roofline_glue, reference_gpu and, with
    the denoise measurement on, the whole-UNet steps/s (full_unet_bench)."""
    _, peak_bw, _ = load_peaks()
    yield "roofline_glue", glue_extra(dev, peak_bw)
    yield "reference_gpu", reference_gpu_extra(dev, cfg_F_mats(1).to(dev))
    if denoise:
        try:
            res = full_unet_bench(dev, max(2, min(steps, 3)))
        except Exception as exc:
            res = {"error": f"{type(exc).__name__}: {exc}"}
        yield "full_unet", res


def run_extras(args):
    """``--impl extras``: the child process of collect_extras -- one JSON line per finished measurement."""
    from cvd_b200 import _lib
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(0)
    _lib.load()
    for key, res in extras_measurements(torch.device("cuda", 0), args.steps, not args.no_denoise):
        print(json.dumps({"extra": key, "result": res}), flush=True)


def collect_extras(cmd, timeout_s, keys):
    """Runs the extras in a CHILD process and returns {key: result}.  They are explanations beside the headline numbers:
    a measurement that dies (a CUDA error is sticky for its whole process) or hangs must not take the line's value / e2e /
    roofline with it, so the child is bounded by ``timeout_s`` and whatever it did not deliver is reported as an error."""
    note = None
    try:
        proc = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, cwd=ROOT)
        stdout = proc.stdout
        if proc.returncode != 0:
            note = f"extras process exited with {proc.returncode}: " + " | ".join((proc.stderr or "").strip().splitlines()[-3:])
    except subprocess.TimeoutExpired as exc:
        stdout = exc.stdout.decode(errors="replace") if isinstance(exc.stdout, bytes) else (exc.stdout or "")
        note = f"extras process killed after {timeout_s} s"
    except OSError as exc:
        stdout, note = "", f"extras process not started: {exc}"
    out = {}
    for ln in stdout.splitlines():
        if ln.startswith('{"extra"'):
            try:
                d = json.loads(ln)
                out[d["extra"]] = d["result"]
            except (ValueError, KeyError):
                pass
    for k in keys:
        out.setdefault(k, {"error": note or "extras process delivered no result"})
    return out


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    import cvd_b200
    import golden_inputs as gi
    from cvd_b200 import _lib, ops, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    # CVD_BENCH_DRYRUN_DEVICE exists for tests/test_bench_dryrun.py only: the CPU build box runs this function's control flow
    # with the kernels and the CUDA runtime replaced by test doubles.  The real ops refuse anything but CUDA tensors.
    dev = torch.device(os.environ.get("CVD_BENCH_DRYRUN_DEVICE", "cuda"), local)
    # one process per GPU: keep this rank's pinned staging buffers and host threads on the GPU's own NUMA node
    prev_affinity = sharding.bind_host_thread_near_gpu(local)
    shard = peer = None
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: if the caller asks for NCCL logging (NCCL_DEBUG=INFO, e.g. to count
        # ranks / see NVLS) it goes to stderr instead of being silenced
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
        shard = sharding.ViewShard()
        if args.exchange in ("peer", "dma"):
            peer = sharding.PeerExchange(shard)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    _lib.load()
    R = R_PER_GPU
    R_glob = R * world
    peak_tf, peak_bw, peak_src = load_peaks()

    # geometry of this rank's rows + partner rows (global numbering when sharded)
    F_rows = cfg_F_mats(1).to(dev)                                 # same pose pair on every rank
    if shard is None:
        partner = torch.cat([torch.arange(R // 2, R), torch.arange(0, R // 2)]).to(torch.int32).to(dev)
    else:
        # view v pairs with the view half the ring away (always on another rank when world > 1)
        glob = (np.arange(R_glob) + R_glob // 2) % R_glob
        partner = torch.from_numpy(shard.local_kv_index(glob, R_glob).astype(np.int32)).to(dev)

    # ---------------- value: kernels only, inputs resident ----------------
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    bufs = []
    for feat, C in SHAPES:
        N, d = feat * feat, C // HEADS
        per_call = []
        for _ in range(CALLS_PER_SHAPE):
            q = torch.randn(R, N, C, device=dev, generator=gen).to(torch.bfloat16)
            if peer is None:
                k = torch.empty(R, N, C, device=dev, dtype=torch.bfloat16)
                vt = torch.empty(R, HEADS, ops.dv_rows(d), N, device=dev, dtype=torch.bfloat16)
                xk = xv = None
            else:   # K / V^T resident in symmetric memory: the peers pull their partner rows from here
                k, pk, hk = peer.alloc((R, N, C), torch.bfloat16, dev)
                vt, pv, hv = peer.alloc((R, HEADS, ops.dv_rows(d), N), torch.bfloat16, dev)
                xk = (pk, hk, [hk.get_buffer(r_, (R, N, C), torch.bfloat16) for r_ in range(world)])
                xv = (pv, hv, [hv.get_buffer(r_, (R, HEADS, ops.dv_rows(d), N), torch.bfloat16) for r_ in range(world)])
            k.copy_(torch.randn(R, N, C, device=dev, generator=gen))
            vt.zero_()
            vt[:, :, :d] = torch.randn(R, HEADS, d, N, device=dev, generator=gen).to(torch.bfloat16)
            vt[:, :, d] = 1.0
            per_call.append((q, k, vt, xk, xv))
        bufs.append(per_call)
    # landing buffers of the pulled partner rows (one pair per shape, reused by its 10 calls)
    land = [None if peer is None else (torch.empty_like(pc[0][1]), torch.empty_like(pc[0][2])) for pc in bufs]
    ident = None if peer is None else peer.identity_index(R, dev)
    theta = torch.tensor([0.5], device=dev)
    big_events = []
    side_k, side_v = torch.cuda.Stream(), torch.cuda.Stream()
    ev_b = torch.cuda.Event()
    ev_k, ev_v = [torch.cuda.Event() for _ in range(8)], [torch.cuda.Event() for _ in range(8)]

    def parse_chunks(env, default):
        spec = os.environ.get(env)
        sizes = [int(x) for x in spec.split(",")] if spec else default
        assert sum(sizes) == R and all(x > 0 for x in sizes) and len(sizes) <= 8, (env, sizes)
        out, lo = [], 0
        for n in sizes:
            out.append((lo, lo + n))
            lo += n
        return out

    # Row chunks of the copy-engine exchange (N > 1): the attention of chunk c runs under the transfer of chunk c + 1.
    # Chunk sizes are whole waves of the attention grid (a launch rounds up to waves: its CTAs wait for the launch in front
    # to complete, programmatic dependent launch only hides their prologue):
    #   32x32, d = 40: 592 CTA slots (four per SM), 64 CTAs per row -> 9.25 rows per wave.  Attention-bound (214 us of
    #     attention against 136 us of transfer): only the FIRST chunk's transfer is exposed, so it is the smallest:
    #     9 | 18 | 37 rows = 1 | 2 | 4 waves, transfers of 19 | 38 | 79 us under attention of 31 | 62 | 124 us.
    #   16x16, d = 80: 296 slots, 16 CTAs per row -> 18.5 rows per wave.  Transfer-bound (68 us against 36 us): the step
    #     ends one chunk of attention after the LAST byte lands, so the last chunks are the small ones: 28 | 18 | 18.
    chunk_plan = {32: parse_chunks("CVD_BENCH_CHUNKS32", [9, 18, 37]), 16: parse_chunks("CVD_BENCH_CHUNKS16", [28, 18, 18])}
    # issued eagerly (graph capture failed or --no-graph) the host is the limiter and every extra chunk costs more than it
    # hides (round 1, measured: three chunks 6.36 ms per step against 5.97 ms for two): the two-chunk plan of round 1
    chunk_plan_eager = {32: parse_chunks("CVD_BENCH_CHUNKS32", [27, 37]), 16: parse_chunks("CVD_BENCH_CHUNKS16", [32, 32])}
    runs_of = run_plan = run_plan_eager = None
    if shard is not None:
        # contiguous runs of the partner map inside [lo, hi): (dst_lo, dst_hi, source rank, source row)
        ph = partner.cpu().numpy()

        def runs_of(lo, hi):
            rr, a_ = [], lo
            for i in range(lo + 1, hi + 1):
                if i == hi or ph[i] != ph[i - 1] + 1 or ph[i] // R != ph[a_] // R:
                    rr.append((a_, i, int(ph[a_]) // R, int(ph[a_]) % R))
                    a_ = i
            return rr

        partner_local = (partner % R).to(torch.int32)      # row inside the source rank's buffer
        run_plan = {f: [runs_of(lo, hi) for lo, hi in ch] for f, ch in chunk_plan.items()}
        run_plan_eager = {f: [runs_of(lo, hi) for lo, hi in ch] for f, ch in chunk_plan_eager.items()}
        runs_whole = runs_of(0, R)

    keep = {"outs": None}      # a list here collects every attention output of a step (graph replay vs eager check)

    def timed_attn(record, *a):
        """One attention launch; with ``record`` its CUDA-event pair joins the event list of the current call (a call of
        the dominant shape is one launch, or one launch per row chunk: the chunk times add up to the call)."""
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            o = ops.epi_attn_fwd(*a)
            e1.record()
            big_events[-1].append((e0, e1))
        else:
            o = ops.epi_attn_fwd(*a)
        if keep["outs"] is not None:
            keep["outs"].append(o)

    def slice_plan(chunks_of, runs_for):
        """Every view the copy-engine exchange needs, cut ONCE: per shape, per call, per row chunk the (landing, peer
        source) pairs of K and V^T and the chunk's q / landing / identity-partner views.  Issued eagerly the step is
        host-bound (round 1), and a tensor slice costs about as much host time as a launch."""
        plan = []
        for (feat, _C), per_call, ld in zip(SHAPES, bufs, land):
            calls = []
            for q, _k, _vt, xk, xv in per_call:
                if feat not in chunks_of:
                    calls.append(None)
                    continue
                chunks = []
                for (lo, hi), runs_c in zip(chunks_of[feat], runs_for[feat]):
                    ck = [(ld[0][dlo:dhi], xk[2][src][slo:slo + (dhi - dlo)]) for (dlo, dhi, src, slo) in runs_c]
                    cv = [(ld[1][dlo:dhi], xv[2][src][slo:slo + (dhi - dlo)]) for (dlo, dhi, src, slo) in runs_c]
                    chunks.append((ck, cv, q[lo:hi], ld[0][lo:hi], ld[1][lo:hi], ident[:hi - lo], lo, hi))
                calls.append(chunks)
            plan.append(calls)
        return plan

    dma = shard is not None and peer is not None and args.exchange == "dma"
    plan_graph = slice_plan(chunk_plan, run_plan) if dma else None
    plan_eager = slice_plan(chunk_plan_eager, run_plan_eager) if dma else None
    in_kernel_src = None        # latents <= 8x8 read the partner's K / V^T in-kernel when it all lives on ONE peer
    if dma and len({rk for (_, _, rk, _) in runs_whole}) == 1:
        in_kernel_src = runs_whole[0][2]

    def value_step(record, eager_plan=False):
        plan = plan_eager if eager_plan else plan_graph
        main_stream = torch.cuda.current_stream()
        for si, ((feat, C), per_call, ld) in enumerate(zip(SHAPES, bufs, land)):
            rec = record and feat == SHAPES[0][0]
            for ci, (q, k, vt, xk, xv) in enumerate(per_call):
                if rec:
                    big_events.append([])
                rowp = ops.epi_row_params(F_rows, feat, F_MAT_SIZE, theta, _lib.FF_RANDOM_SLOPE)
                if shard is None:
                    timed_attn(rec, q, k, vt, rowp, partner, HEADS, F_MAT_SIZE)
                elif peer is None:            # baseline: NCCL all-gather of every rank's K and V^T
                    timed_attn(rec, q, shard.gather_rows(k), shard.gather_rows(vt), rowp, partner, HEADS, F_MAT_SIZE)
                elif args.exchange == "dma":
                    # one cross-rank barrier (the partner's K / V^T of this call are complete) ...
                    sharding.barrier(xk[1])
                    if feat * feat <= 64 and in_kernel_src is not None:
                        # ... then, one query tile per (row, head): the kernel reads the partner's K / V^T tiles straight
                        # out of the peer's HBM (unicast TMA over NVLink, plain loads at 4x4) -- every byte crosses once,
                        # no landing buffer, no copy launch
                        timed_attn(False, q, xk[2][in_kernel_src], xv[2][in_kernel_src], rowp, partner_local, HEADS, F_MAT_SIZE)
                        continue
                    # ... then the partner rows cross NVLink on the COPY ENGINES in row chunks (K and V^T on two side
                    # streams), each chunk's attention running under the transfer of the next (no SMs spent on the pull)
                    chunks = plan[si][ci]
                    ev_b.record(main_stream)
                    side_k.wait_event(ev_b)
                    side_v.wait_event(ev_b)
                    with torch.cuda.stream(side_k):
                        for c, ch in enumerate(chunks):
                            for dst, src_ in ch[0]:
                                dst.copy_(src_, non_blocking=True)
                            ev_k[c].record(side_k)
                    with torch.cuda.stream(side_v):
                        for c, ch in enumerate(chunks):
                            for dst, src_ in ch[1]:
                                dst.copy_(src_, non_blocking=True)
                            ev_v[c].record(side_v)
                    for c, (_ck, _cv, q_c, lk_c, lv_c, id_c, lo, hi) in enumerate(chunks):
                        main_stream.wait_event(ev_k[c])
                        main_stream.wait_event(ev_v[c])
                        timed_attn(rec, q_c, lk_c, lv_c, rowp[lo:hi], id_c, HEADS, F_MAT_SIZE)
                else:                         # one cross-rank barrier, then pull only the partner rows over NVLink (SM kernel)
                    ks = peer.pull(xk[0], xk[1], partner, ld[0], rows_per_src=R)
                    vs = ops.peer_gather_rows(xv[0], partner, ld[1], rows_per_src=R)
                    timed_attn(rec, q, ks, vs, rowp, ident, HEADS, F_MAT_SIZE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Phase A: the step issued EAGERLY -- at N > 1 with the two-chunk plan of round 1, the path that has run on hardware.
    # (The CUDA-graph replay of the N > 1 step is measured in phase B, after every other number of the line is safe.)
    for _ in range(max(args.warmup, 3)):
        value_step(False, eager_plan=True)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def time_steps(step_fn):
        """EXACTLY args.steps steps between two events, barrier + synchronize on both sides, clocks sampled meanwhile."""
        clocks = ClockSampler(local)
        clocks.start()
        t0.record()
        for _ in range(args.steps):
            step_fn()
        t1.record()
        barrier()
        return t0.elapsed_time(t1) / args.steps, clocks.stop()

    def call_ms():      # the dominant kernel per call of R rows (a call is one launch, or one per row chunk: they add up)
        return float(np.mean([sum(a.elapsed_time(b) for a, b in call) for call in big_events if call]))

    launches0 = ops.LAUNCHES
    ms_value, clk = time_steps(lambda: value_step(True, eager_plan=True))
    launches = ops.LAUNCHES - launches0
    big_ms = call_ms()
    exchange_issue = "eager" if world > 1 else None

    # every attention shape on its own (N = 1): the 10 calls of the shape back to back on their distinct buffers
    by_shape = None
    if world == 1 and not args.no_extras:
        try:
            by_shape = []
            for (feat, C), per_call in zip(SHAPES, bufs):
                rowp = ops.epi_row_params(F_rows, feat, F_MAT_SIZE, theta, _lib.FF_RANDOM_SLOPE)
                for q, k, vt, _xk, _xv in per_call[:2]:
                    ops.epi_attn_fwd(q, k, vt, rowp, partner, HEADS, F_MAT_SIZE)
                torch.cuda.synchronize()
                a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                for q, k, vt, _xk, _xv in per_call:
                    ops.epi_attn_fwd(q, k, vt, rowp, partner, HEADS, F_MAT_SIZE)
                b_.record()
                torch.cuda.synchronize()
                us = a_.elapsed_time(b_) / len(per_call) * 1e3
                fl = 4.0 * R * (feat * feat) ** 2 * C
                by_shape.append({"N": feat * feat, "C": C, "head_dim": C // HEADS, "us_per_launch": us,
                                 "tflops": fl / us / 1e6, "frac": fl / us / 1e6 / peak_tf,
                                 "kernel": "epi_attn_small_kernel (mma.sync, warp-level)" if feat <= 4 else
                                           f"epi_attn_fwd_kernel<{C // HEADS},{feat}> (tcgen05)"})
        except Exception as exc:
            by_shape = {"error": f"{type(exc).__name__}: {exc}"}

    # ---------------- e2e: layer API with host buffers ----------------
    layers, hosts = [], []
    for feat, C in SHAPES:
        attn = cvd_b200.EpiSelfAttention(attention_mode="Epi_Self", cross_attention_dim=None, query_dim=C, heads=HEADS,
                                         dim_head=C // HEADS, epi_position_encoding=True,
                                         epi_position_encoding_feat_max_size=64,
                                         epi_position_encoding_F_mat_size=F_MAT_SIZE, epi_rand_slope_ff=True)
        attn.load_state_dict(gi.attn_state_dict(C, 10 + feat), strict=False)
        attn = attn.to(dev).eval()
        attn.processor.shard = shard
        attn.processor.peer = peer
        layers.append(attn)
        x_host = torch.randn(R, feat * feat, C, generator=torch.Generator().manual_seed(feat)).pin_memory()
        y_host = torch.empty_like(x_host).pin_memory()
        hosts.append((x_host, y_host))
    kv_index = None if shard is None else partner.cpu().numpy().astype(np.int64)
    h2d = sum(CALLS_PER_SHAPE * x.numel() * 4 for x, _ in hosts)
    d2h = h2d

    # H2D, compute and D2H run on three streams with a 3-deep ring of device buffers per shape, so the
    # PCIe transfers of neighbouring calls overlap the kernels (all of it inside the timed region).
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    ring = 3
    dev_in = [[torch.empty_like(x, device=dev) for _ in range(ring)] for x, _ in hosts]
    ev_in = [[torch.cuda.Event() for _ in range(ring)] for _ in hosts]
    ev_done = [[torch.cuda.Event() for _ in range(ring)] for _ in hosts]
    main = torch.cuda.current_stream()

    def e2e_step():
        with torch.no_grad():
            for li, (attn, (x_host, y_host)) in enumerate(zip(layers, hosts)):
                for c in range(CALLS_PER_SHAPE):
                    b = c % ring
                    with torch.cuda.stream(s_in):
                        s_in.wait_event(ev_done[li][b])          # the previous user of this slot has consumed it
                        dev_in[li][b].copy_(x_host, non_blocking=True)
                        ev_in[li][b].record(s_in)
                    main.wait_event(ev_in[li][b])
                    y, _aux = attn(dev_in[li][b], F_mats=F_rows if kv_index is None else [F_rows, kv_index])
                    ev_done[li][b].record(main)
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(ev_done[li][b])
                        y_host.copy_(y, non_blocking=True)
                        y.record_stream(s_out)
        main.wait_stream(s_out)

    for _ in range(3):
        e2e_step()
    barrier()
    e_steps = max(3, min(args.steps, 10))
    t0.record()
    for _ in range(e_steps):
        e2e_step()
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1) / e_steps

    # release what the later measurements do not need (phase B at N > 1 replays the kernel step: its buffers stay)
    del layers, hosts, dev_in
    if world == 1:
        del bufs, land
    torch.cuda.empty_cache()
    ms_denoise, views, ms_eager, by_views, config4 = (None, 0, None, None, None) if args.no_denoise else \
        denoise_bench(dev, world, rank, shard, peer, max(3, min(args.steps, 5)))

    glue_rows = ref_gpu = full_unet = None
    if world == 1 and not args.no_extras:
        keys = EXTRA_KEYS if not args.no_denoise else EXTRA_KEYS[:2]
        if "CVD_BENCH_DRYRUN_DEVICE" in os.environ:      # the dry run's test doubles live in THIS process
            ex = dict(extras_measurements(dev, args.steps, not args.no_denoise))
        else:
            ex = collect_extras([sys.executable, os.path.abspath(__file__), "--impl", "extras", "--steps", str(args.steps)]
                                + (["--no-denoise"] if args.no_denoise else []), EXTRAS_TIMEOUT_S, keys)
        glue_rows, ref_gpu, full_unet = ex.get("roofline_glue"), ex.get("reference_gpu"), ex.get("full_unet")

    # max over ranks
    if world > 1:
        tt = torch.tensor([ms_value, ms_e2e, big_ms, ms_denoise or 0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_value, ms_e2e, big_ms, md = (float(v) for v in tt.tolist())
        ms_denoise = md if ms_denoise is not None else None
    flops = step_flops(R_glob)
    big_flops = 4.0 * R * 1024 * 1024 * 320
    achieved = big_flops / (big_ms * 1e-3) / 1e12

    line = None
    sharding.restore_host_affinity(prev_affinity)      # the CPU baseline gets every core the box has
    if rank == 0:
        cb_times, cb_flops, cores = cpu_reference_step(R, reps=8, warm=2)      # ~10 s of host work on 16 cores
        cb_val = cb_flops / min(cb_times) / 1e12
        line = {
            "metric": "epi_attn_tflops", "value": flops / (ms_value * 1e-3) / 1e12, "unit": "TFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": flops / (ms_e2e * 1e-3) / 1e12, "unit": "TFLOP/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "cvd_b200.EpiSelfAttention.forward (fp32 pinned host tensors; H2D / compute / D2H on 3 streams)"},
            "gpu_launches": launches,
            "clocks": clk,
            "roofline": {"bound": "tensor", "kernel": "epi_attn_fwd_kernel<40,32> (R=64, N=M=1024, C=320)",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": ncu_traffic_bytes(), "traffic_unit": "bytes per launch (ncu, profiles/r01_epi_attn_d40_ncu_full_v7.json)",
                         "peak_source": peak_src, "ms_per_launch": big_ms,
                         "algorithmic_flops_per_launch": big_flops},
            "cpu_baseline": {"value": cb_val, "unit": "TFLOP/s", "cores": cores, "kind": "port",
                             "sample": f"oracle EpiSelfAttention.forward, 1 call per shape ({cb_flops / 1e9:.1f} GFLOP per pass), 2 warm + 8 timed passes, best pass"},
        }
        if exchange_issue is not None:
            line["config"]["exchange_issue"] = exchange_issue
        if by_shape is not None:
            line["roofline_by_shape"] = by_shape
        if glue_rows is not None:
            line["roofline_glue"] = glue_rows
        if ref_gpu is not None:
            line["reference_gpu"] = ref_gpu
        if ms_denoise is not None:
            line["denoise"] = {"steps_per_s": 1e3 / ms_denoise, "ms_per_step": ms_denoise, "views": views,
                               "rows_per_gpu": R_PER_GPU, "guidance_scale": 8.5,
                               "launch": "one CUDA graph per UNet forward" if world == 1 else "eager", "ms_per_step_eager": ms_eager,
                               "unet": "hot-path skeleton cvd_b200.sync_unet.SyncPathUNet (20 temporal + 20 epi modules at "
                                       "256x256; SD-1.5 spatial blocks excluded -- not comparable to a full-UNet number)",
                               "loop": "sampler.sample_pair" if world == 1 else "sampler.sample_multiview (view-sharded)"}
            if by_views is not None:
                line["denoise"]["by_views"] = by_views
            if full_unet is not None:
                line["denoise"]["full_unet"] = full_unet
    if world == 1:
        print(json.dumps(line), flush=True)
        return

    # Phase B (N > 1): two measurements whose code had not run on hardware when this was written -- the step replayed as ONE
    # CUDA graph, and the unit-sharded config-4 loop.  Rank 0 HOLDS the finished line (phase A numbers) meanwhile; each
    # result is merged into it in one piece; an exception (a CUDA error is sticky for the process) or LATE_TIMEOUT_S without
    # an end prints the line as it stands and leaves the process without touching CUDA again.
    lock = threading.Lock()
    done = [False]

    def emit(note=None, hard=False):
        with lock:
            if done[0]:
                return
            done[0] = True
            if rank == 0:
                if note:
                    line["config"]["late_measurements"] = note
                print(json.dumps(line), flush=True)
        if hard:
            sys.stdout.flush()
            sys.stderr.flush()
            hard_exit(0)

    def merge(fn):
        with lock:
            if rank == 0 and not done[0]:
                fn(line)

    timer = threading.Timer(LATE_TIMEOUT_S, emit, kwargs=dict(
        note=f"graph replay / config-4 measurement still running after {LATE_TIMEOUT_S} s: line printed without them", hard=True))
    timer.daemon = True
    timer.start()
    try:
        if not args.no_graph and args.exchange == "dma":
            # The step is ~15 stream operations per call (barrier, event hand-offs, peer copies, chunk launches): issued
            # eagerly the host cannot keep up with the GPU (round 1: SMs idle, 385 W).  Captured ONCE (device-side barriers,
            # copy-engine transfers and kernels alike, wave-sized chunk plan) and replayed, the host is out of the loop.
            graph, why, graph_outs, launches_per_step = None, None, None, 0
            l0 = ops.LAUNCHES
            try:
                cap = torch.cuda.Stream()
                cap.wait_stream(torch.cuda.current_stream())
                g_ = torch.cuda.CUDAGraph()
                keep["outs"] = []
                # thread_local: the process group's watchdog thread polls CUDA events while we capture; only THIS thread's
                # calls are part of (and may invalidate) the capture
                with torch.cuda.graph(g_, stream=cap, capture_error_mode="thread_local"):
                    value_step(False)
                graph_outs = keep["outs"]
                launches_per_step = ops.LAUNCHES - l0
                graph = g_
            except Exception as exc:      # capture is an optimisation: the eager numbers stand, and the line says why
                why = f"capture failed on rank {rank} ({type(exc).__name__}: {exc})"
                print(f"[bench] {why}", file=sys.stderr, flush=True)
            keep["outs"] = None
            # Every rank must take the same path, and must agree on it BEFORE the first replay: a replay executes the step's
            # cross-rank barriers, which a rank whose capture failed would never answer (capturing executes nothing).
            flag = torch.tensor([1 if graph is not None else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                graph, why = None, why or "capture failed on another rank"
            if graph is not None:
                for _ in range(2):
                    graph.replay()
                barrier()
                # the replay must give the eager step's bits (same kernels on the same rows; only chunking and issue differ)
                keep["outs"] = []
                value_step(False, eager_plan=True)
                eager_outs, keep["outs"] = keep["outs"], None
                torch.cuda.synchronize()

                def by_shape_rows(outs_):      # consecutive outputs of one (N, C) -> [calls * R, N, C], rows in call order
                    groups = []
                    for o in outs_:
                        if groups and groups[-1][0].shape[1:] == o.shape[1:]:
                            groups[-1].append(o)
                        else:
                            groups.append([o])
                    return [torch.cat(g) for g in groups]
                ga, ea = by_shape_rows(graph_outs), by_shape_rows(eager_outs)
                same = len(ga) == len(ea) and all(x.shape == y.shape and torch.equal(x, y) for x, y in zip(ga, ea))
                del ga, ea, eager_outs
                flag.fill_(1 if same else 0)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                if int(flag.item()) == 0:
                    graph, why = None, "the graph replay's outputs differ from the eager step's: not used"
            if graph is not None:
                ms_graph, clk_graph = time_steps(graph.replay)
                big_events.clear()      # the dominant kernel's launch times under the graph's chunk plan: two eager passes
                for _ in range(2):
                    value_step(True)
                barrier()
                tg = torch.tensor([ms_graph, call_ms()], device=dev, dtype=torch.float64)
                dist.all_reduce(tg, op=dist.ReduceOp.MAX)
                ms_graph, big_graph = (float(v) for v in tg.tolist())

                def use_graph(ln):
                    ln["config"]["exchange_ms_per_step"] = {"eager": ln["ms_per_step"], "graph": ms_graph,
                                                            "graph_equals_eager_bitwise": True}
                    if ms_graph < ln["ms_per_step"]:
                        ach = big_flops / (big_graph * 1e-3) / 1e12
                        ln.update(value=flops / (ms_graph * 1e-3) / 1e12, ms_per_step=ms_graph, clocks=clk_graph,
                                  gpu_launches=launches_per_step * args.steps)
                        ln["config"]["exchange_issue"] = "one CUDA graph per step"
                        ln["roofline"].update(achieved=ach, frac=ach / peak_tf, ms_per_launch=big_graph)
                merge(use_graph)
            else:
                merge(lambda ln: ln["config"].__setitem__("graph", why))
        if config4 is not None:
            res4 = config4()
            merge(lambda ln: ln["denoise"].setdefault("by_views", {}).update(res4))
    except Exception as exc:      # most likely a CUDA error: this process can print, and nothing else
        timer.cancel()
        emit(note=f"ended by {type(exc).__name__}: {str(exc).splitlines()[0] if str(exc) else ''}", hard=True)
        return
    timer.cancel()
    emit()
    dist.barrier()      # the other ranks start tearing NCCL down (and logging it, under NCCL_DEBUG=INFO) after the line is out
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "extras"])
    ap.add_argument("--no-denoise", action="store_true", help="skip the UNet-level steps/s measurement")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip roofline_by_shape / roofline_glue / reference_gpu (N = 1 only; a few seconds)")
    ap.add_argument("--exchange", default="dma", choices=["peer", "dma", "allgather"],
                    help="N>1: partner rows over NVLink by the copy engines in row chunks under the attention, small grids read "
                         "in-kernel (dma, default), by our pull kernel (peer), or NCCL all-gather (baseline)")
    ap.add_argument("--no-graph", action="store_true", help="N>1: skip phase B's CUDA-graph replay of the step (the eagerly issued step is timed either way)")
    args = ap.parse_args()
    global EXCHANGE
    if args.exchange == "allgather":
        EXCHANGE = "NCCL all-gather"
    if args.exchange == "peer":
        EXCHANGE = "NVLink peer pull, cvd_peer_gather_rows"
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "extras":
        run_extras(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
