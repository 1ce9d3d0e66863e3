"""CPU dry run of ``bench.py``'s N = 1 control flow (the driver-facing artifact: one JSON line per run), with the kernels
replaced by the plain-torch test double (tests/emulated_ops.py), the CUDA runtime calls (streams, events, graphs, pinned
memory) by no-op stand-ins, and the workload shrunk to two small shapes.  It cannot say anything about speed; it checks
that every code path of the default run executes and that the line carries every key the contract names -- the parts of
``bench.py`` that changed in round 2 after GPU access was closed (value loop refactor, per-shape roofline, glue roofline,
reference-GPU leg) would otherwise run for the first time in the driver's hands."""
import contextlib
import io
import json
import os
import sys
import threading
import types

import pytest
import torch

import emulated_ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Stream:
    cuda_stream = 0

    def wait_stream(self, s):
        pass

    def wait_event(self, e):
        pass


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def wait(self, stream=None):
        pass

    ms = 1.0                 # what every event pair "measures"; a stand-in graph replay makes it 0.25 (see _Graph)

    def elapsed_time(self, other):
        return _Event.ms


class _Graph:
    replays = 0
    modes = []

    on_replay = None
    replay_ms = None         # when set: after a replay every event pair "measures" this (a replayed step that is faster)

    def replay(self):
        _Graph.replays += 1
        if _Graph.replay_ms is not None:
            _Event.ms = _Graph.replay_ms
        if _Graph.on_replay is not None:
            _Graph.on_replay()


CAPTURE_FAILS = False        # the stand-in capture raises at its end, as an invalidated capture does


@contextlib.contextmanager
def _graph_ctx(graph, stream=None, capture_error_mode="global"):
    _Graph.modes.append(capture_error_mode)
    yield
    if CAPTURE_FAILS:
        raise RuntimeError("CUDA error: operation failed due to a previous error during capture")


def fake_cuda(monkeypatch):
    cur = _Stream()
    monkeypatch.setattr(_Event, "ms", 1.0)
    monkeypatch.setattr(_Graph, "on_replay", None)
    monkeypatch.setattr(_Graph, "replay_ms", None)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: cur)
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "CUDAGraph", _Graph)
    monkeypatch.setattr(torch.cuda, "graph", _graph_ctx)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None)


def test_bench_default_run_control_flow(monkeypatch):
    fake_cuda(monkeypatch)
    emulated_ops.install(monkeypatch)
    from cvd_b200 import _lib, ops
    monkeypatch.setattr(_lib, "load", lambda: None)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setenv("CVD_BENCH_DRYRUN_DEVICE", "cpu")
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "SHAPES", [(8, 320), (4, 320)])
    monkeypatch.setattr(bench, "CALLS_PER_SHAPE", 2)
    # the glue measurement at its real shapes is minutes of CPU time: keep its code path, shrink its shapes
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import profile_glue
    real_measure = profile_glue.measure
    monkeypatch.setattr(profile_glue, "measure", lambda dev, peak: real_measure(dev, peak, shapes=((4, 64),), iters=2))
    monkeypatch.setattr(bench, "cpu_reference_step", lambda R, reps, warm: ([1.0] * reps, 1e9, 1))
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=1, impl="ours", no_denoise=True, no_extras=False, exchange="dma",
                                 no_graph=False)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        bench.run_ours(args)
    lines = [ln for ln in out.getvalue().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline",
                "roofline_by_shape", "roofline_glue", "reference_gpu"):
        assert key in d, key
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["dtype"] == "bf16"
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert isinstance(d["roofline_by_shape"], list) and len(d["roofline_by_shape"]) == 2, d["roofline_by_shape"]
    assert isinstance(d["roofline_glue"], list) and len(d["roofline_glue"]) >= 10, d["roofline_glue"]
    assert "error" not in d["reference_gpu"], d["reference_gpu"]
    assert len(d["reference_gpu"]["per_shape"]) == 2 and all(r["rel_l2_vs_fp32"] < 3e-2 for r in d["reference_gpu"]["per_shape"])
    # 2 shapes x 2 calls x (prepass + attention) per step
    assert d["gpu_launches"] == 2 * 2 * 2 * 2


def test_bench_reference_arm_honours_warmup(monkeypatch, capsys):
    sys.path.insert(0, ROOT)
    import bench
    seen = {}

    def fake(R, reps, warm):
        seen["warm"], seen["reps"] = warm, reps
        return [0.5] * reps, 1e9, 3
    monkeypatch.setattr(bench, "cpu_reference_step", fake)
    bench.run_reference(types.SimpleNamespace(gpus=1, steps=4, warmup=2))
    d = json.loads(capsys.readouterr().out.strip())
    assert seen == {"warm": 2, "reps": 4} and d["impl"] == "reference" and d["warmup"] == 2 and d["steps"] == 4
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


@pytest.mark.parametrize("no_graph", [False, True, "capture_fails", "peer_capture_fails", "replay_dies", "replay_hangs"])
def test_bench_sharded_run_control_flow(monkeypatch, no_graph):
    """The N > 1 branch of ``run_ours`` in ONE process: rank 0 of a pretended world of 2 whose "peer" buffers alias its own
    (symmetric memory, the process group and the collectives replaced by stand-ins).  Exercises the chunk / run plans of
    the copy-engine exchange, the in-kernel branch of the small latents, the graph-capture path with its launch
    accounting, the eager passes that time the dominant kernel, the sharded e2e layers and the JSON line."""
    import numpy as np
    import torch.distributed as dist
    fake_cuda(monkeypatch)
    emulated_ops.install(monkeypatch)
    from cvd_b200 import _lib, ops, sharding
    monkeypatch.setattr(_lib, "load", lambda: None)
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setenv("CVD_BENCH_DRYRUN_DEVICE", "cpu")
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setattr(dist, "init_process_group", lambda *a, **k: None)
    monkeypatch.setattr(dist, "destroy_process_group", lambda *a, **k: None)
    monkeypatch.setattr(dist, "get_world_size", lambda group=None: 2)
    monkeypatch.setattr(dist, "get_rank", lambda group=None: 0)
    monkeypatch.setattr(dist, "barrier", lambda *a, **k: None)
    # "peer_capture_fails": the MIN all-reduce of the captured-flag comes back 0 although this rank's capture worked
    scenario = no_graph
    monkeypatch.setattr(dist, "all_reduce", lambda t, *a, **k: t.zero_() if (scenario == "peer_capture_fails" and k.get("op") == dist.ReduceOp.MIN) else None)
    monkeypatch.setattr(sys.modules[__name__], "CAPTURE_FAILS", no_graph == "capture_fails")
    _Graph.replays, _Graph.modes = 0, []
    failed = no_graph in ("capture_fails", "peer_capture_fails")
    late = no_graph in ("replay_dies", "replay_hangs")
    no_graph = no_graph is True or failed or late
    barriers = []

    class Handle:
        def __init__(self, t):
            self.t = t

        def get_buffer(self, rank, shape, dtype):
            return self.t                                   # the "peer's" copy aliases the local one

        def barrier(self, channel=0, timeout_ms=0):
            barriers.append(timeout_ms)

    class FakePeer:
        mode = "hidden"

        def __init__(self, shard):
            self.shard = shard

        def alloc(self, shape, dtype, device):
            t = torch.empty(*shape, dtype=dtype, device=device)
            return t, [0, 0], Handle(t)

        def identity_index(self, n, device):
            return torch.arange(n, device=device, dtype=torch.int32)

        def exchange(self, x, partner):
            return x[(partner.long() % x.shape[0])]

    monkeypatch.setattr(sharding, "PeerExchange", FakePeer)
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "SHAPES", [(32, 16), (16, 16), (8, 64), (4, 64)])      # all four exchange branches, tiny C
    monkeypatch.setattr(bench, "HEADS", 2)
    monkeypatch.setattr(bench, "CALLS_PER_SHAPE", 1)
    monkeypatch.setattr(bench, "cpu_reference_step", lambda R, reps, warm: ([1.0] * reps, 1e9, 1))
    args = types.SimpleNamespace(gpus=2, steps=2, warmup=1, impl="ours", no_denoise=True, no_extras=False, exchange="dma",
                                 no_graph=scenario is True)
    exits = []

    def fake_exit(code):
        exits.append(code)
        if threading.current_thread() is threading.main_thread():      # (the timer thread just ends)
            raise SystemExit(code)
    monkeypatch.setattr(bench, "hard_exit", fake_exit)
    monkeypatch.setattr(_Graph, "replay_ms", 0.25)
    if scenario == "replay_dies":
        def boom():
            raise RuntimeError("CUDA error: unspecified launch failure\nCompile with TORCH_USE_CUDA_DSA ...")
        monkeypatch.setattr(_Graph, "on_replay", boom)
    if scenario == "replay_hangs":
        import time
        monkeypatch.setattr(bench, "LATE_TIMEOUT_S", 0.3)
        monkeypatch.setattr(_Graph, "on_replay", lambda: time.sleep(0.8) if _Graph.replays == 1 else None)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        try:
            bench.run_ours(args)
        except SystemExit:
            assert scenario == "replay_dies"
    lines = [ln for ln in out.getvalue().splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                   # ONE line, whatever happened to the late measurements
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["config"]["views_total"] == 4 and "copy engines" in d["config"]["parallelism"]
    assert all(t_ == sharding.BARRIER_TIMEOUT_MS for t_ in barriers) and len(barriers) > 0
    # captured beside a live process-group watchdog thread: only this thread's calls belong to the capture
    assert _Graph.modes == ([] if scenario is True else ["thread_local"])
    # Phase A always times the eagerly issued step (the two-chunk plan of round 1): per step 4 prepasses + 2 + 2 chunk
    # launches of 32x32 / 16x16 + 1 + 1 in-kernel launches; the chunk launches of 1.0 "ms" each add up to one call.
    eager = {"launches": 2 * (4 + 2 * 2 + 1 + 1), "ms_per_launch": 2.0, "ms_per_step": 0.5}
    if scenario is False:
        # phase B: the replay (3 + 3 wave-sized chunks) gave the eager step's bits and was faster -> it is the line's value
        assert _Graph.replays == 2 + 2
        assert d["config"]["exchange_issue"] == "one CUDA graph per step"
        assert d["config"]["exchange_ms_per_step"] == {"eager": 0.5, "graph": 0.125, "graph_equals_eager_bitwise": True}
        assert d["gpu_launches"] == 2 * (4 + 2 * 3 + 1 + 1) and d["ms_per_step"] == 0.125
        assert d["roofline"]["ms_per_launch"] == 3 * 0.25
    else:
        assert d["config"]["exchange_issue"] == "eager"
        assert d["gpu_launches"] == eager["launches"] and d["ms_per_step"] == eager["ms_per_step"]
        assert d["roofline"]["ms_per_launch"] == eager["ms_per_launch"]
        # a rank never replays (= executes cross-rank barriers) unless EVERY rank captured
        assert _Graph.replays == {True: 0, "capture_fails": 0, "peer_capture_fails": 0, "replay_dies": 1, "replay_hangs": 4}[scenario]
        if scenario == "capture_fails":
            assert "capture failed on rank 0" in d["config"]["graph"]
        if scenario == "peer_capture_fails":
            assert d["config"]["graph"] == "capture failed on another rank"
        if scenario == "replay_dies":                        # the held line goes out, then the process leaves
            assert exits == [0] and "CUDA error: unspecified launch failure" in d["config"]["late_measurements"]
        if scenario == "replay_hangs":                       # the timer printed the held line
            assert exits == [0] and "still running after" in d["config"]["late_measurements"]
    assert "roofline_by_shape" not in d and "reference_gpu" not in d      # N = 1 extras only


@pytest.mark.parametrize("world", [1, 2])
def test_denoise_bench_control_flow(monkeypatch, world):
    """``bench.denoise_bench`` at reduced size (4x4 latents, one UNet level): N = 1 -- the graphed two-video loop and the
    2 / 4 / 6-view triple; N = 2 (one process pretending to be rank 0 of 2, peers aliased) -- the view-sharded loop and the
    unit-sharded config-4 loop (6 views x 2 x 2 = 24 units, 12 per rank), each with its guarded error reporting."""
    import numpy as np
    import torch.distributed as dist
    fake_cuda(monkeypatch)
    emulated_ops.install(monkeypatch)
    from cvd_b200 import ops
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "DENOISE_FEAT", 4)
    monkeypatch.setattr(bench, "DENOISE_LEVELS", 1)
    shard = peer = None
    if world > 1:
        monkeypatch.setattr(dist, "get_world_size", lambda group=None: world)
        monkeypatch.setattr(dist, "get_rank", lambda group=None: 0)
        monkeypatch.setattr(dist, "barrier", lambda *a, **k: None)
        monkeypatch.setattr(dist, "all_reduce", lambda t, *a, **k: None)
        from cvd_b200 import sharding
        shard = sharding.ViewShard()

        class AliasPeer:
            mode = "hidden"

            def exchange(self, x, partner):
                return x[(partner.long() % x.shape[0])].contiguous()

            def identity_index(self, n, device):
                return torch.arange(n, device=device, dtype=torch.int32)
        peer = AliasPeer()
    ms, views, ms_eager, by_views, config4 = bench.denoise_bench(torch.device("cpu"), world, 0, shard, peer, 2)
    assert ms == 0.5 and views == 2 * world            # fake events: 1.0 "ms" over 2 steps
    if world == 1:
        assert ms_eager == 0.5 and set(by_views) == {"2", "4", "6"} and by_views["6"]["rows"] == 192 and config4 is None
    else:
        assert by_views is None
        by_views = config4()          # run last by the caller (bench.run_ours, phase B)
        assert "error" not in by_views["config4"], by_views["config4"]
        assert by_views["config4"]["units_per_gpu"] == 12 and by_views["config4"]["views"] == 6 and by_views["config4"]["finite"]


def test_full_unet_bench_control_flow(monkeypatch):
    """``bench.full_unet_bench`` (whole-UNet denoise steps/s + the reference-GPU twin beside it) at the reduced width of the
    UNet fixtures: graphed loop, twin construction with oracle-backed modules on the same state dict, the difference
    between the two noise predictions (bf16-emulated sync path vs the fp32 oracle: a few 1e-3)."""
    import golden_inputs as gi
    fake_cuda(monkeypatch)
    emulated_ops.install(monkeypatch)
    from cvd_b200 import ops
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    sys.path.insert(0, ROOT)
    import bench
    kw = dict(gi.UNET_KWARGS, amp=False)             # fp32 SD blocks on the CPU (bf16 autocast convolutions are slow there)
    monkeypatch.setattr(bench, "FULL_UNET_KWARGS", kw)
    monkeypatch.setattr(bench, "FULL_UNET_TEXT_LEN", 7)
    out = bench.full_unet_bench(torch.device("cpu"), 2)
    assert out["steps_per_s"] == 2000.0 and out["views"] == 2 and out["parameters_M"] > 1.0
    assert out["forward"]["rel_l2"] < 2e-2, out["forward"]


def test_extras_run_in_a_bounded_child_process(tmp_path):
    """collect_extras: results the child printed are kept; whatever it did not deliver (crash, non-zero exit, hang past the
    limit, nothing printed) becomes an {"error": ...} entry -- the bench line survives either way."""
    sys.path.insert(0, ROOT)
    import bench
    keys = ("roofline_glue", "reference_gpu", "full_unet")
    ok = ('import json; print("noise"); print(json.dumps({"extra": "roofline_glue", "result": [1, 2]}), flush=True); '
          'print(json.dumps({"extra": "reference_gpu", "result": {"speedup_vs_fp32": 9.0}}), flush=True); ')
    out = bench.collect_extras([sys.executable, "-c", ok + 'print(json.dumps({"extra": "full_unet", "result": {"steps_per_s": 3.0}}))'],
                               30, keys)
    assert out == {"roofline_glue": [1, 2], "reference_gpu": {"speedup_vs_fp32": 9.0}, "full_unet": {"steps_per_s": 3.0}}
    # dies after two results (a sticky CUDA error ends in an abort): the two survive, the third says why it is missing
    out = bench.collect_extras([sys.executable, "-c", ok + 'import os, sys; sys.stderr.write("CUDA error: boom\\n"); os._exit(134)'],
                               30, keys)
    assert out["roofline_glue"] == [1, 2] and "134" in out["full_unet"]["error"] and "boom" in out["full_unet"]["error"]
    # hangs: killed at the limit, delivered results kept
    out = bench.collect_extras([sys.executable, "-c", ok + 'import time; time.sleep(60)'], 3, keys)
    assert out["reference_gpu"] == {"speedup_vs_fp32": 9.0} and "killed after 3 s" in out["full_unet"]["error"]
    out = bench.collect_extras([str(tmp_path / "no_such_binary")], 3, keys)
    assert all("not started" in out[k]["error"] for k in keys)


def test_extras_child_refuses_to_run_without_a_gpu():
    """The real ``bench.py --impl extras`` child on this box (no GPU): it must fail loudly (no CPU fallback), and the
    parent-side collector turns that into error entries."""
    if torch.cuda.is_available():
        pytest.skip("CPU build box only")
    sys.path.insert(0, ROOT)
    import bench
    out = bench.collect_extras([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "extras", "--steps", "2"], 300,
                               bench.EXTRA_KEYS)
    assert all("error" in out[k] and "needs a GPU" in out[k]["error"] for k in bench.EXTRA_KEYS), out
