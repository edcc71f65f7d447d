"""N>1 path on CPU: world_size-2 (and 3) gloo runs of the slab / halo algorithm model against the single-domain oracle."""
import pathlib
import subprocess
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("inplace", [0, 2, 4])
@pytest.mark.parametrize("world", [2, 3])
def test_slab_halo_model_gloo(world, inplace):
    """inplace = 0: two population buffers per rank; S >= 2: every rank streams IN PLACE (one buffer, launches of S - 1 planes in dependence
    order, rank r with shift S + r) and the halo moves between the current frames of the ranks."""
    import os
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(29700 + world + 10 * inplace), str(ROOT / "tests" / "dist_slab_model.py")], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, SLAB_MODEL_INPLACE=str(inplace)))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bit-identical=True" in r.stdout
