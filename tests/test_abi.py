"""The C-ABI library builds for sm_100a, loads without a GPU and exports every symbol include/fg_api.h declares
(no compute calls here).  Also: the product path refuses to run without CUDA instead of falling back."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "fg_api.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fg_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    names = declared_functions()
    for must in ["fg_project_fwd", "fg_project_bwd", "fg_isect_emit", "fg_radix_sort_pairs_u64_u32",
                 "fg_isect_offsets", "fg_rasterize_fwd", "fg_rasterize_bwd", "fg_knn_f32", "fg_last_error"]:
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in fg_api.h but not exported"
    lib.fg_abi_version.restype = ctypes.c_int
    from freegaussian_b200 import _lib
    assert lib.fg_abi_version() == _lib.ABI_VERSION


def test_python_binding_matches_header(built_lib):
    from freegaussian_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_functions()
    _lib.lib()  # binds argtypes for every symbol


def test_library_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", str(built_lib)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_bad_arguments_return_error_codes_not_crashes(built_lib):
    from freegaussian_b200 import _lib
    L = _lib.lib()
    rc = L.fg_rasterize_fwd(1, 10, 99, 64, 64, 16, None, None, None, None, None, None, 0, 99, -1, 0, None, None, 0,
                            None, None, None, None, None)
    assert rc == 1 and b"CH" in L.fg_last_error()
    rc = L.fg_project_fwd(0, 10, None, None, None, None, None, 64, 64, 0.3, 0.01, 1e10, 0.0, 16, -1, 0, None, None,
                          None, None, 0, None, None, None, None, None, None, 0, -1, -1, -1, None, None, None)
    assert rc == 1
    with pytest.raises(AssertionError):
        _lib.check(rc)


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    from freegaussian_b200.knn import k_nearest
    from freegaussian_b200.rendering import rasterization
    with pytest.raises(RuntimeError, match="no CPU path"):
        rasterization(torch.zeros(4, 3), torch.ones(4, 4), torch.ones(4, 3), torch.ones(4), torch.zeros(4, 16, 3),
                      torch.eye(4)[None], torch.eye(3)[None], 32, 32, sh_degree=3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        k_nearest(torch.rand(10, 3), 3)


def test_boundary_asserts_like_gsplat():
    from freegaussian_b200.rendering import rasterization
    ok = dict(means=torch.zeros(4, 3), quats=torch.ones(4, 4), scales=torch.ones(4, 3), opacities=torch.ones(4),
              colors=torch.zeros(4, 16, 3), viewmats=torch.eye(4)[None], Ks=torch.eye(3)[None], width=32, height=32)
    with pytest.raises(AssertionError):
        rasterization(**{**ok, "means": torch.zeros(4, 2)}, sh_degree=3)
    with pytest.raises(AssertionError):
        rasterization(**ok, sh_degree=3, render_mode="BGR")
    with pytest.raises(AssertionError):
        rasterization(**ok, sh_degree=5)
    with pytest.raises(AssertionError):
        rasterization(**{**ok, "colors": torch.zeros(4, 4, 3)}, sh_degree=3)  # too few SH bases
    with pytest.raises(NotImplementedError):
        rasterization(**ok, sh_degree=3, camera_model="fisheye")


def test_compat_helpers():
    from freegaussian_b200.compat import get_viewmat, num_sh_bases, quat_to_rotmat
    from oracle.render import quat_to_rotmat as oracle_q2r
    assert [num_sh_bases(d) for d in range(4)] == [1, 4, 9, 16]
    q = torch.randn(7, 4)
    assert torch.allclose(quat_to_rotmat(q), oracle_q2r(q), atol=1e-6)
    R = quat_to_rotmat(torch.randn(3, 4))
    c2w = torch.cat([R, torch.randn(3, 3, 1)], -1)
    vm = get_viewmat(c2w)
    # world->camera of the camera centre is the origin; y/z are flipped (utils.py:162-179)
    centre = torch.cat([c2w[:, :, 3], torch.ones(3, 1)], -1)
    assert torch.allclose(torch.einsum("cij,cj->ci", vm, centre)[:, :3], torch.zeros(3, 3), atol=1e-5)
    assert torch.allclose(vm[:, :3, :3], (R * torch.tensor([1.0, -1.0, -1.0])).transpose(1, 2), atol=1e-6)


def test_every_kernel_waits_for_its_stream_predecessor():
    """All launches are programmatic dependent launches (csrc/common.cuh FG_LAUNCH): a kernel that did not start
    with pdl_wait() could run ahead of the kernel that produces its input."""
    n = 0
    for path in sorted((ROOT / "freegaussian_b200" / "csrc").glob("*.cu")):
        text = path.read_text()
        assert "<<<" not in text, f"{path.name}: raw launch bypasses FG_LAUNCH"
        for m in re.finditer(r"__global__[^{;]*\{", text):
            body = text[m.end():m.end() + 1200]
            first = [ln.strip() for ln in body.split("\n") if ln.strip() and not ln.strip().startswith("//")]
            # declarations of shared memory may precede it; nothing that touches memory may
            head = []
            for ln in first:
                head.append(ln)
                if ln.startswith("pdl_wait();"):
                    break
            assert head and head[-1].startswith("pdl_wait();"), f"{path.name}: kernel without pdl_wait: {first[:2]}"
            for ln in head[:-1]:
                assert ln.startswith(("extern __shared__", "__shared__", "using ", "constexpr ")), (path.name, ln)
            n += 1
    assert n >= 30


def test_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of every struct of fg_api.h as gcc lays them out == the ctypes mirrors in _lib.py."""
    from freegaussian_b200 import _lib

    structs = {"fg_adam_segment": _lib.AdamSegment, "fg_refine_config": _lib.RefineConfig, "fg_refine_array": _lib.RefineArray,
               "fg_mlp_pack_segment": _lib.MlpPackSegment}
    header_fields = {"fg_refine_array": {"in_": "in"}}  # ctypes cannot name a field `in`
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT / "include" / "fg_api.h"}"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            hname = header_fields.get(cname, {}).get(fname, fname)
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {hname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    for line in out.strip().splitlines():
        cname, field, value = line.split()
        cls = structs[cname]
        want = ctypes.sizeof(cls) if field == "size" else getattr(cls, field).offset
        assert int(value) == want, line


def test_network_entry_points_validate_arguments(built_lib):
    """The tensor-core entry points reject bad shapes / NULL operands with an error code before touching the device."""
    from freegaussian_b200 import _lib
    L = _lib.lib()
    assert L.fg_mlp_linear(_lib.MLP_RELU, 128, 256, None, 33, None, 0, None, None, None, None, None, None, None) == 1
    assert b"multiples of 32" in L.fg_last_error()
    assert L.fg_mlp_linear(_lib.MLP_RELU, 128, 256, None, 256, None, 0, None, None, None, None, None, None, None) == 1
    assert b"NULL" in L.fg_last_error()
    assert L.fg_mlp_linear(_lib.MLP_RELU, 0, 256, None, 256, None, 0, None, None, None, None, None, None, None) == 0  # empty input
    assert L.fg_mlp_wgrad(16, None, None, 256, None, 256, 0, None, None) == 1
    assert L.fg_mlp_wgrad(16, None, None, 256, None, 255, 0, None, None) == 1 and b"aligned" in L.fg_last_error()
    assert L.fg_mlp_wgrad(0, None, None, 256, None, 256, 0, None, None) == 0
    assert L.fg_deform_embed(10, None, None, None, 40, 10, 96, None, None) == 1 and b"wider" in L.fg_last_error()
    assert L.fg_mlp_pack(99, None, None) == 1


def test_oracle_is_test_infrastructure_only():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import the oracle; the product and tools/ never do."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b", re.M)
    offenders = []
    for path in list((ROOT / "freegaussian_b200").rglob("*.py")) + list((ROOT / "tools").rglob("*.py")):
        if pat.search(path.read_text()):
            offenders.append(str(path.relative_to(ROOT)))
    assert not offenders, offenders
    bench = (ROOT / "bench.py").read_text()
    body = bench[bench.index("def cpu_arm("):bench.index("def run_reference(")]
    assert len(pat.findall(bench)) == len(pat.findall(body)) > 0  # every oracle import of bench.py sits inside cpu_arm()
    entry = (ROOT / "__graft_entry__.py").read_text()
    assert len(pat.findall(entry)) == len(pat.findall(entry[entry.index("def smoke("):]))  # and inside smoke()
