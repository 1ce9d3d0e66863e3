"""-m gpu, needs >= 2 GPUs (skipped otherwise): the view-sharded paths on real NVLink peers.  Every test launches
``torchrun`` on a worker under ``tests/multigpu/`` (one process per GPU, NCCL + symmetric memory, rendezvous on
127.0.0.1) and requires its success line; the workers compare sharded results with the un-sharded computation of the
same rows (or with a local copy of the peer's rows) and assert inside every rank."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(worker, nproc, ok_line, timeout=480):
    if not torch.cuda.is_available() or torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs, this box has {torch.cuda.device_count() if torch.cuda.is_available() else 0}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multigpu", worker)]
    # own session: on a timeout the launcher AND its workers are ended (the process group started here, nothing else)
    proc = subprocess.Popen(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = proc.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        import signal
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        pytest.fail(f"{worker} did not finish within {timeout} s\n" + out[-2000:] + "\n---- stderr ----\n" + err[-2000:])
    tail = out[-3000:] + "\n---- stderr ----\n" + err[-3000:]
    assert proc.returncode == 0 and ok_line in out, tail
    return out


def test_layer_peer_exchange_equals_unsharded():
    """EpiSelfAttention on 2 ranks (partner rows of the LayerNorm-ed hidden states pulled over NVLink, and the NCCL
    all-gather baseline) == the rows of the un-sharded layer."""
    _torchrun("mg_peer_exchange.py", 2, "PEER EXCHANGE OK")


def test_attention_reads_peer_kv_in_kernel():
    """The attention kernels reading the partner's K / V^T straight out of the peer's HBM (unicast TMA over NVLink at 8x8,
    plain loads at 4x4; never the cluster-multicast kernels) are bit-identical to the same call on
    a local copy."""
    _torchrun("mg_peer_direct.py", 2, "DIRECT PEER READ OK")


def test_sharded_sampler_equals_unsharded():
    """V-view DDIM loop with views sharded over 2 ranks == the views of the un-sharded loop (pairing, device F
    recompute, partner rows over NVLink, re-noise drawn full-size and sliced)."""
    _torchrun("mg_sharded_sampler.py", 2, "SHARDED SAMPLER OK")
