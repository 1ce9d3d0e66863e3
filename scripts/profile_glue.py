"""Achieved HBM bandwidth of every glue kernel of libcvd_b200.so at the BASELINE config-2 shapes
(R = 64 rows = 4 videos x 16 frames), CUDA events, against MEASURED_PEAKS.json's copy bandwidth.

    python scripts/profile_glue.py [--json out.json]

Bytes are the ALGORITHMIC bytes of DESIGN.md section 4 (compulsory reads + writes, nothing for
re-reads).  Every kernel rotates over enough distinct buffers that a pass touches > 2x the 126 MB L2
between reuses of one buffer.  ``measure()`` is also what bench.py calls for its ``roofline_glue`` key.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import torch

L2 = 126e6


def measure(dev, peak, shapes=((32, 320), (16, 640), (8, 1280)), iters=20, verbose=False, eager=False):
    """Returns a list of dicts {kernel, shape, bytes, us, gbs, frac} (frac = achieved / ``peak`` GB/s)."""
    from cvd_b200 import _lib, ops
    rows = []

    def bench(name, shape_note, nbytes, make, run):
        """make() -> one argument set; run(args) launches the kernel once."""
        n_sets = max(2, int(2.5 * L2 // max(nbytes, 1)) + 1)
        n_sets = min(n_sets, 24)
        sets = [make() for _ in range(n_sets)]
        for s in sets[:2]:
            run(s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if eager:     # plain launches (for ncu); timings are then launch-rate bound
            e0.record()
            for i in range(iters):
                run(sets[i % n_sets])
            e1.record()
        else:
            # the launches replay from a CUDA graph so that the small shapes are not bounded by the Python call rate
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                with torch.cuda.graph(graph, stream=side):
                    for i in range(iters):
                        run(sets[i % n_sets])
            graph.replay()
            torch.cuda.synchronize()
            e0.record()
            graph.replay()
            e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        gbs = nbytes / us / 1e3
        rows.append(dict(kernel=name, shape=shape_note, bytes=nbytes, us=us, gbs=gbs, frac=gbs / peak))
        if verbose:
            print(f"{name:28s} {shape_note:34s} {nbytes / 1e6:8.1f} MB {us:8.1f} us {gbs:8.0f} GB/s  "
                  f"{100 * gbs / peak:5.1f} % of measured copy peak")
        del sets

    B, Fr = 4, 16
    for feat, C in shapes:
        HW = feat * feat
        T = B * Fr * HW            # tokens
        note = f"R=64 N={HW} C={C}"
        g = torch.ones(C, device=dev)
        b = torch.zeros(C, device=dev)

        # GroupNorm + layout: video fp32 read (statistics pass re-reads it out of L2) -> bf16 tokens
        bench("groupnorm_tokens (epi)", note, T * C * (4 + 2),
              lambda: torch.randn(B, C, Fr, feat, feat, device=dev),
              lambda x: ops.groupnorm_tokens(x, g, b, 32, 1e-6, _lib.LAYOUT_BF_HW))
        bench("groupnorm_tokens (temporal)", note, T * C * (4 + 2),
              lambda: torch.randn(B, C, Fr, feat, feat, device=dev),
              lambda x: ops.groupnorm_tokens(x, g, b, 32, 1e-6, _lib.LAYOUT_BHW_F))
        # LayerNorm: fp32 in, bf16 out
        bench("layernorm", note, T * C * (4 + 2),
              lambda: torch.randn(B * Fr, HW, C, device=dev),
              lambda x: ops.layernorm(x, g, b))
        # residual update + LayerNorm in one pass: fp32 stream read + written, bf16 branch read, bf16 normed written
        bench("residual_layernorm", note, T * C * (4 + 2 + 4 + 2),
              lambda: (torch.randn(B * Fr, HW, C, device=dev), torch.randn(B * Fr, HW, C, device=dev, dtype=torch.bfloat16)),
              lambda a: ops.residual_layernorm_(a[0], a[1], b, 1.0, g, b))
        # GEGLU: bf16 [T, 8C] -> bf16 [T, 4C]
        bench("geglu", note, T * C * (8 + 4) * 2,
              lambda: torch.randn(B * Fr, HW, 8 * C, device=dev, dtype=torch.bfloat16),
              lambda x: ops.geglu(x))
        # residual: fp32 stream (read + write) + bf16 branch (+ bf16 copy out)
        bench("residual_add", note, T * C * (4 + 2 + 4),
              lambda: (torch.randn(B * Fr, HW, C, device=dev), torch.randn(B * Fr, HW, C, device=dev, dtype=torch.bfloat16)),
              lambda a: ops.residual_add_(a[0], a[1], b, 1.0, False))
        bench("residual_add (+bf16 copy)", note, T * C * (4 + 2 + 4 + 2),
              lambda: (torch.randn(B * Fr, HW, C, device=dev), torch.randn(B * Fr, HW, C, device=dev, dtype=torch.bfloat16)),
              lambda a: ops.residual_add_(a[0], a[1], b, 1.0, True))
        # tokens -> video + residual: bf16 tokens + fp32 residual read, fp32 video written
        bench("tokens_to_video (epi)", note, T * C * (2 + 4 + 4),
              lambda: (torch.randn(B * Fr, HW, C, device=dev, dtype=torch.bfloat16), torch.randn(B, C, Fr, feat, feat, device=dev)),
              lambda a: ops.tokens_to_video(a[0], b, a[1], _lib.LAYOUT_BF_HW))
        bench("tokens_to_video (temporal)", note, T * C * (2 + 4 + 4),
              lambda: (torch.randn(B * HW, Fr, C, device=dev, dtype=torch.bfloat16), torch.randn(B, C, Fr, feat, feat, device=dev)),
              lambda a: ops.tokens_to_video(a[0], b, a[1], _lib.LAYOUT_BHW_F))
        # temporal attention: q, k, v read once, out written once (bf16)
        bench("temporal_attn", note, T * C * 4 * 2,
              lambda: tuple(torch.randn(B * HW, Fr, C, device=dev, dtype=torch.bfloat16) for _ in range(3)),
              lambda a: ops.temporal_attn(a[0], a[1], a[2], 8))
        bench("add_pose_tokens", note, T * C * (2 + 4 + 2),
              lambda: (torch.randn(B * HW, Fr, C, device=dev, dtype=torch.bfloat16), torch.randn(B, C, Fr, feat, feat, device=dev)),
              lambda a: ops.add_pose_tokens(a[0], a[1]))
        bench("scale_add", note, T * C * 3 * 2,
              lambda: (torch.randn(T, C, device=dev, dtype=torch.bfloat16), torch.randn(T, C, device=dev, dtype=torch.bfloat16)),
              lambda a: ops.scale_add(a[0], a[1], 1.0))
        if verbose:
            print()

    # reference point measured the same way: a plain device copy of the res-1 fp32 stream
    x = [torch.randn(B * Fr * 1024 * 320, device=dev) for _ in range(6)]
    y = torch.empty_like(x[0])
    bench("torch copy_ (reference point)", "84 MB fp32 -> fp32", x[0].numel() * 8, lambda: None, lambda _: y.copy_(x[len(rows) % 6]))
    return rows


if __name__ == "__main__":
    peak = 6547.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    out_rows = measure(torch.device("cuda:0"), peak, verbose=True, eager=os.environ.get("PROFILE_GLUE_EAGER") == "1")
    if "--json" in sys.argv:
        json.dump({"peak_gbs": peak, "rows": out_rows}, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
