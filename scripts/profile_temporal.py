import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, cvd_b200, golden_inputs as gi
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
feat, C = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 320)
t = cvd_b200.get_motion_module(C, "Vanilla", dict(num_attention_heads=8, num_transformer_block=1, attention_block_types=["Temporal_Self", "Temporal_Self"],
                               temporal_position_encoding=True, temporal_position_encoding_max_len=32, zero_initialize=False))
blk = t.temporal_transformer.transformer_blocks[0]
blk.attention_blocks[0].set_processor(cvd_b200.PoseAdaptorAttnProcessor(hidden_size=C, pose_feature_dim=C, query_condition=True, key_value_condition=True))
t.load_state_dict(gi.temporal_state_dict(C, 9), strict=False); t = t.to(dev).eval()
x = torch.randn(4, C, 16, feat, feat, device=dev); pose = torch.randn(4, C, 16, feat, feat, device=dev)
with torch.no_grad():
    for _ in range(3): t(x, cross_attention_kwargs={"pose_feature": pose})
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): t(x, cross_attention_kwargs={"pose_feature": pose})
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))
