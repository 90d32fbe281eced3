"""Data parallelism, world_size 2 over gloo on CPU (kernels through the emulator): sharding the batch over two ranks with the
asynchronous per-stack gradient all-reduce and the fused optimiser reproduces the single-process step on the whole batch —
SUM of gradients for `m_iwae` (a sum over the batch), SUM / world for `elbo` (a mean); `negInfoNCE` gathers the projections so
that its negatives are the global batch."""
import os
import socket
import subprocess
import sys

import pytest
import torch

import dp_worker

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("objective", ["m_iwae", "elbo", "contrast", "accumulate"])
def test_two_ranks_match_one(emu, tmp_path, objective):
    out = str(tmp_path / "dp.pt")
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "dp_worker.py"), objective, out], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(l[-3000:] for l in logs)
    got = torch.load(out, weights_only=False)
    # single process, whole batch
    model = dp_worker.build(objective)
    x, us = dp_worker.data(4)
    loss, grads, params = dp_worker.step(model, x, us, objective, average=False)
    assert abs(got["loss"] - loss) < 1e-5 * max(1.0, abs(loss)), (got["loss"], loss)
    gmax = max(float(g.abs().max()) for g in grads.values())
    for n, g in grads.items():
        w = got["grads"][n] * (0.5 if objective == "elbo" else 1.0)       # the elbo optimiser divides the SUM by the world size
        assert float((w - g).abs().max()) <= 2e-5 * float(g.abs().max()) + 1e-5 * gmax + 5e-7, (n, float((w - g).abs().max()))
    # updated parameters: entries whose gradient is analytically zero (key biases: softmax is shift-invariant) hold round-off
    # that Adam normalises to +-lr, so only entries with a real gradient are compared
    for n, p in params.items():
        real = grads[n].abs() > 1e-6 if n in grads else torch.zeros_like(p, dtype=torch.bool)
        d = (got["params"][n] - p).abs()
        assert float(d[real].max() if real.any() else 0.0) < 2e-5, (n, float(d.max()))
        assert float(d.max()) <= 2.1e-2                       # never more than two learning-rate steps
