"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's own PyTorch GPU path for one EpiSelfAttention call.

The oracle restatement of the reference moved to the GPU unchanged: fp32 eager (optionally under bf16 autocast, the
fairness variant BASELINE.md section 3 asks for), dense [R,N,N] bias generated per call, the NaN check with its host
sync, the mask replicated over the 8 heads, ``F.scaled_dot_product_attention`` with that mask
(epi_module.py:293-329,395-407; attention_processor.py:585-619).  Used by tests/test_gpu_speedup.py and by the
``reference_gpu`` key of bench.py as the thing the product is timed BESIDE; nothing under cvd_b200/ imports it."""
import torch
import torch.nn.functional as F

from . import epi_oracle as O


def reference_gpu_layer(x, sd, heads, F_mats, S=256, autocast=False):
    """EpiSelfAttention.forward as the reference runs it on a GPU: get_attn_map -> NaN scrub -> mask replicated over
    the heads (prepare_attention_mask, attention_processor.py:585-587) -> to_q/k/v -> SDPA(mask) -> to_out.
    x [R,N,C] fp32 on the GPU, sd: fp32 state dict of the layer on the same device."""
    R, N, C = x.shape
    with torch.device(x.device):
        bias = O.epi_bias(F_mats, int(N ** 0.5), S)                       # [R,N,N] fp32, horizontal first-frame rule
    if bool(torch.isnan(bias).any()):                                     # the reference's host sync (:404)
        bias = torch.nan_to_num(bias)
    mask = bias.repeat_interleave(heads, dim=0).view(R, heads, N, N)      # x8 copy, as prepare_attention_mask does
    idx = torch.cat([torch.arange(R // 2, R), torch.arange(0, R // 2)]).to(x.device)
    d = C // heads
    with torch.autocast(x.device.type, dtype=torch.bfloat16, enabled=autocast):
        enc = x[idx]
        q = F.linear(x, sd["to_q.weight"]).view(R, N, heads, d).transpose(1, 2)
        k = F.linear(enc, sd["to_k.weight"]).view(R, N, heads, d).transpose(1, 2)
        v = F.linear(enc, sd["to_v.weight"]).view(R, N, heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=mask.to(q.dtype), dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(R, N, C)
        return F.linear(o, sd["to_out.0.weight"], sd["to_out.0.bias"]).float()
