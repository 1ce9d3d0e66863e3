"""Time one EpiModule / VanillaTemporalModule forward per UNet resolution and break it down by kernel
(torch profiler).  R = 64 rows (2 views x 2 cfg x 16 frames), 256x256 images.

    python scripts/profile_module.py [--table]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np, torch
import cvd_b200, golden_inputs as gi

dev = torch.device("cuda:0")
Fm = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "geometry.npz"))["F_mats"])
F4 = torch.cat([Fm[:16], Fm[:16], Fm[16:], Fm[16:]]).reshape(4, 16, 3, 3).to(dev)

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

tot_epi = tot_tmp = 0.0
for feat, C, n_mod in ((32, 320, 5), (16, 640, 5), (8, 1280, 5), (4, 1280, 5)):
    m = cvd_b200.get_epi_module(C, dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Epi_Self", "Epi_Self"],
                                        epi_position_encoding=True, epi_position_encoding_feat_max_size=64,
                                        epi_position_encoding_F_mat_size=256, epi_rand_slope_ff=True, zero_initialize=True))
    m.load_state_dict(gi.module_state_dict(C, 7), strict=False); m = m.to(dev).eval()
    t = cvd_b200.get_motion_module(C, "Vanilla", dict(num_attention_heads=8, num_transformer_block=1,
                                   attention_block_types=["Temporal_Self", "Temporal_Self"], temporal_position_encoding=True,
                                   temporal_position_encoding_max_len=32, zero_initialize=False))
    blk = t.temporal_transformer.transformer_blocks[0]
    blk.attention_blocks[0].set_processor(cvd_b200.PoseAdaptorAttnProcessor(hidden_size=C, pose_feature_dim=C, query_condition=True, key_value_condition=True))
    t.load_state_dict(gi.temporal_state_dict(C, 9), strict=False); t = t.to(dev).eval()
    x = torch.randn(4, C, 16, feat, feat, device=dev)
    pose = torch.randn(4, C, 16, feat, feat, device=dev)
    gm, gt = cvd_b200.GraphedModule(m), cvd_b200.GraphedModule(t)
    with torch.no_grad():
        te = timeit(lambda: m(x, F_mats=F4))
        tt = timeit(lambda: t(x, cross_attention_kwargs={"pose_feature": pose}))
        ye, _ = m(x, F_mats=F4.clone())
        torch.manual_seed(5); ye, _ = m(x, F_mats=F4); ye = ye.clone()
        torch.manual_seed(5); yg, _ = gm(x, F_mats=F4)      # capture (consumes warm-up draws first)
        ge = timeit(lambda: gm(x, F_mats=F4))
        gtt = timeit(lambda: gt(x, cross_attention_kwargs={"pose_feature": pose}))
        yt = t(x, cross_attention_kwargs={"pose_feature": pose}); ytg = gt(x, cross_attention_kwargs={"pose_feature": pose})
    print(f"feat={feat:2d} C={C:4d}: EpiModule {te*1e3:8.1f} us (graph {ge*1e3:8.1f})   VanillaTemporalModule {tt*1e3:8.1f} us (graph {gtt*1e3:8.1f})   "
          f"temporal graph==eager: {torch.equal(yt, ytg)}")
    te, tt = ge, gtt
    tot_epi += n_mod * te; tot_tmp += n_mod * tt
    if "--table" in sys.argv and feat == 32:
        from torch.profiler import profile, ProfilerActivity
        with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            for _ in range(5): m(x, F_mats=F4)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
print(f"per UNet forward (20 + 20 modules): epi {tot_epi:.2f} ms, temporal {tot_tmp:.2f} ms")
