"""Hardware multi-GPU parity (SURVEY.md section 4 tier 4; the reference's only data-parallel hook is the sampler,
datasets/loader/dataset_sampler.py:94-103): needs >= 2 GPUs, skipped otherwise.  One process per GPU through
torch.distributed.run on 127.0.0.1; the checks live in tests/mgpu_worker.py."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(mode, tmp_path, world=2):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    out = str(tmp_path / ("%s.json" % mode))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "mgpu_worker.py"), mode, out]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, "worker failed:\n%s\n%s" % (res.stdout[-3000:], res.stderr[-3000:])
    with open(out) as f:
        return json.load(f)


def test_sharded_inference_is_bit_identical(tmp_path):
    s = _run("infer", tmp_path)
    print(s)
    assert s["ok"] and all(s["levels_bit_identical"])


def test_allreduced_gradients_equal_single_process(tmp_path):
    s = _run("train", tmp_path)
    print(s)
    assert s["ok"] and s["max_rel_l2_vs_single_process"] <= 1e-3 and s["max_rel_l2_accumulated"] <= 1e-3
    assert s["buckets"] >= 4 and s["allreduce_mb_per_step"] > 50
