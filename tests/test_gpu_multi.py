"""Multi-GPU (x-slab ring over NCCL) parity: needs >= 2 visible GPUs, otherwise skipped (the 1-GPU box runs the rest)."""
import pathlib
import subprocess
import sys

import pytest

import biblbm_b200 as bib

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world", [2, 4])
def test_slab_ring_bit_identical_to_single_domain(world):
    if bib.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", str(ROOT / "tests" / "mgpu" / "slab_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("bit-identical=True") == 3


@pytest.mark.parametrize("world", [2, 4])
def test_multislab_immersed_boundary(world):
    """Ghost particles, partial-sum exchange and ownership hand-off across slab faces and the periodic seam."""
    if bib.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29613", str(ROOT / "tests" / "mgpu" / "ib_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "phase_equal=True" in r.stdout


@pytest.mark.parametrize("world", [2, 4])
def test_multislab_channel_suspension(world):
    """BASELINE configs[3] in miniature (rbcs_in_channel: wall planes, ParallelPlatesBoundary, body force, a dense row of RBCs crossing every
    slab face) sharded over `world` GPUs against the undivided lattice on one GPU."""
    if bib.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import os
    env = dict(os.environ, IB_CHECK_CASE="channel", IB_CHECK_STEPS="120")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29615", str(ROOT / "tests" / "mgpu" / "ib_check.py")], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "case=channel" in r.stdout and "phase_equal=True" in r.stdout
