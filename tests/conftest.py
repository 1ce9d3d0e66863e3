import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    d = os.path.join(ROOT, "tests", "golden")
    return {n: np.load(os.path.join(d, n + ".npz")) for n in ("geometry", "bias", "attention", "module", "temporal")}


# ---- CVD_TEST_EMULATE=1: dry run of the -m gpu tests on a box without a GPU -------------------------------------------------
EMULATE = os.environ.get("CVD_TEST_EMULATE") == "1"
# what cannot be dry-run: raw accumulator dumps of the tcgen05 kernel, wall-clock speed assertions, checks that CPU tensors are
# refused (the dry run pretends they are CUDA tensors), peer pulls, CUDA-graph replay (the stand-in graph does not replay),
# and cases whose fp32 reference needs tens of GB / many minutes on the CPU
_NO_DRYRUN = ("test_qk_accumulator_dump", "test_layer_vs_reference_gpu_path", "test_denoise_step_vs_reference_gpu_path",
              "test_epi_module_train_step_config5_shape", "test_attention_forward_backward_config5_shape", "16-64-40",
              "test_cpu_tensor_is_refused", "test_training_requires_cuda", "test_peer_gather_rows_single_gpu",
              "test_graphed_unet_equals_eager", "test_graphed_multiview_loop_equals_eager")


@pytest.fixture(autouse=True)
def _emulate_kernels(request, monkeypatch):
    if not EMULATE or request.node.get_closest_marker("gpu") is None:
        yield
        return
    if any(k in request.node.nodeid for k in _NO_DRYRUN):
        pytest.skip("not part of the CPU dry run")
    import contextlib
    import torch
    import emulated_ops
    emulated_ops.install(monkeypatch, strict_shapes=True)
    from cvd_b200 import ops
    monkeypatch.setattr(ops, "_stream", lambda: 0)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: type("S", (), {"cuda_stream": 0, "wait_stream": lambda self, s: None})())
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: type("S", (), {"wait_stream": lambda self, s: None})())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())

    class _Graph:
        def replay(self):
            pass

    @contextlib.contextmanager
    def _graph_ctx(graph, stream=None):
        yield
    monkeypatch.setattr(torch.cuda, "CUDAGraph", _Graph)
    monkeypatch.setattr(torch.cuda, "graph", _graph_ctx)
    yield
