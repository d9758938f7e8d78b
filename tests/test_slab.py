"""Slab mode: host logic with gloo world_size 2 on the CPU, the collective device solve on >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "slab_worker.py")


def _torchrun(nproc, port, *args, timeout=600, worker=WORKER):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), worker, *args]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("world", [2, 3])
def test_slab_host_logic_gloo(world):
    r = _torchrun(world, 29541 + world, "--host-only")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "slab host logic ok" in r.stdout


@pytest.mark.parametrize("world", [2, 3])
def test_slab_multilevel_host_logic_gloo(world):
    r = _torchrun(world, 29546 + world, "--host-only", worker=os.path.join(ROOT, "tests", "slab_ml_worker.py"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "slab multilevel host logic ok" in r.stdout


@pytest.mark.gpu
def test_slab_solve_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    r = _torchrun(2, 29551, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "max|T_slab - T_single|" in r.stdout
    assert "max|Ti_slab - Ti_single|" in r.stdout          # pfem_interpolate_field in slab mode
    assert "max|Tl_slab - T_slab|" in r.stdout
    assert "max|T021_slab - T021_single|" in r.stdout
    assert "max|Tm_slab - Tm_single|" in r.stdout
    assert "max|Tb_slab - Tb_single|" in r.stdout
    assert "max|V_slab - V_single|" in r.stdout
    assert "max|V201_slab - V201_single|" in r.stdout
    assert "max|Td_slab - Td_single|" in r.stdout          # Dynamic3D in slab mode      # vertical-major mesh: the host cuts a lateral axis


@pytest.mark.gpu
def test_slab_multilevel_two_gpus():
    """the multilevel preconditioner in slab mode = the single-GPU hierarchy (same iteration counts), Static3D and Shockley3D"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    r = _torchrun(2, 29557, timeout=300, worker=os.path.join(ROOT, "tests", "slab_ml_worker.py"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "max|Tml_slab - Tml_single|" in r.stdout and "max|Vml_slab - Vml_single|" in r.stdout and "slab multilevel ok" in r.stdout
