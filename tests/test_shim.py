"""The ``gsplat`` import shim (CPU): the reference's own import statements, verbatim, resolve to this package."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

# freegaussian/freegaussian_model.py:15-21, freegaussian_control_model.py:7-10, preprocess/knn_gaussian.py:9 -- verbatim
REFERENCE_IMPORTS = '''
from gsplat.cuda_legacy._torch_impl import quat_to_rotmat

try:
    from gsplat.rendering import rasterization
except ImportError:
    print("Please install gsplat>=1.0.0")
from gsplat.cuda_legacy._wrapper import num_sh_bases
'''


def test_reference_import_lines_resolve_to_this_package():
    code = REFERENCE_IMPORTS + '''
import torch, freegaussian_b200.rendering as R, freegaussian_b200.compat as C
assert rasterization is R.rasterization and quat_to_rotmat is C.quat_to_rotmat and num_sh_bases is C.num_sh_bases
assert num_sh_bases(3) == 16                                                   # freegaussian_model.py:165
rot = quat_to_rotmat(torch.tensor([[2.0, 0.0, 0.0, 0.0]]))                     # :535 (normalises inside)
assert torch.allclose(rot, torch.eye(3)[None])
try:  # no GPU here: the call must refuse, not fall back to a CPU renderer
    z = torch.zeros
    rasterization(z(4, 3), z(4, 4), z(4, 3), z(4), z(4, 16, 3), torch.eye(4)[None], torch.eye(3)[None], 32, 32, sh_degree=3)
    raise SystemExit("rendered on the CPU?")
except RuntimeError as e:
    assert "no CPU path" in str(e)
print("SHIM-OK")
'''
    env = {"PYTHONPATH": f"{ROOT / 'shims'}:{ROOT}", "PATH": "/usr/bin:/bin"}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and "SHIM-OK" in out.stdout, out.stdout + out.stderr
