"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "facedeform_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from facedeform_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build libfacedeform_gpu.so first (__graft_entry__.build())"
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in facedeform_gpu.h but not exported"
    assert sorted(_lib.EXPORTS) == declared
    assert L.fd_abi_version() == 3


def test_params_defaults_and_clamps_match_the_sop():
    """defaults = SOP_FaceDeform.cpp:117-137, clamps = :249-257 (no GPU needed)."""
    from facedeform_b200 import make_params
    p = make_params()
    assert (p.model, p.term, p.kernel, p.layers, p.maxedges) == (0, 0, 0, 4, 4)
    assert (p.qcoef, p.zcoef, p.radius, p.falloffradius, p.falloffrate) == (1.0, 5.0, 1.0, 1.0, 1.0)
    assert p.lambda_ == pytest.approx(0.1) and tuple(p.weightrange) == (0.0, 1.0)
    assert (p.tangent, p.morphspace, p.doclampweight, p.dofalloff) == (0, 0, 0, 0)
    q = make_params(clamp=True, qcoef=0.0, zcoef=0.01, radius=0.0, layers=-3, maxedges=0, **{"lambda": 0.0})
    assert (q.qcoef, q.zcoef, q.radius, q.lambda_) == pytest.approx((0.1, 0.1, 0.01, 0.01))
    assert (q.layers, q.maxedges) == (1, 1)


def test_params_struct_layout_matches_oracle_semantics(oracle):
    """the oracle's and the product's parameter structs are separate definitions of the same SOP surface."""
    from facedeform_b200 import make_params
    p, o = make_params(), oracle.make_params()
    for f in ("model", "term", "kernel", "qcoef", "zcoef", "radius", "layers", "lambda_", "tangent", "maxedges",
              "dofalloff", "falloffradius", "falloffrate"):
        assert getattr(p, f) == getattr(o, f), f


def test_status_strings_quote_the_reference_messages():
    from facedeform_b200 import _lib
    L = _lib.load()
    assert L.fd_status_string(2) == b"Rest and deform geometry should match."   # SOP_FaceDeform.cpp:232
    assert L.fd_status_string(3) == b"Can't build RBF model."                   # :338
    assert L.fd_status_string(4) == b"Can't solve the problem."                 # :366
    assert L.fd_status_string(7) == b"Can't capture geometry with a rig!"       # :319


def test_no_gpu_fails_loudly():
    """without a GPU fd_ctx_create must fail (FD_E_CUDA), never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from facedeform_b200 import Context, FdError
    with pytest.raises(FdError) as e:
        Context()
    assert e.value.status == 5


def test_header_is_plain_c_and_struct_layouts_match_the_ctypes_mirror(tmp_path):
    """include/facedeform_gpu.h compiles as C (no C++ / torch types in the boundary) and fd_params / fd_report have the
    size and field offsets the ctypes mirror (facedeform_b200/_lib.py) assumes."""
    import shutil
    import subprocess
    from facedeform_b200._lib import FdParams, FdReport
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    src = tmp_path / "layout.c"
    fields_p = [f[0] for f in FdParams._fields_]
    fields_r = [f[0] for f in FdReport._fields_]
    cname = lambda f: "lambda" if f == "lambda_" else f
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "facedeform_gpu.h"', "int main(void) {",
             '  printf("%zu %zu\\n", sizeof(fd_params), sizeof(fd_report));']
    for f in fields_p:
        lines.append(f'  printf("p {f} %zu\\n", offsetof(fd_params, {cname(f)}));')
    for f in fields_r:
        lines.append(f'  printf("r {f} %zu\\n", offsetof(fd_report, {f}));')
    lines += ["  return 0;", "}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    sp, sr = (int(x) for x in out[0].split())
    assert sp == ctypes.sizeof(FdParams) and sr == ctypes.sizeof(FdReport)
    for line in out[1:]:
        if not line:
            continue
        kind, name, off = line.split()
        struct = FdParams if kind == "p" else FdReport
        assert getattr(struct, name).offset == int(off), (kind, name)


def test_fit_equal_separates_fit_from_epilogue_parameters():
    """fd_params_fit_equal: what FaceDeformOp's refit decision compares (the reference refits every cook, :331-363)."""
    import ctypes as C
    from facedeform_b200 import _lib, make_params
    L = _lib.load()
    eq = lambda a, b: L.fd_params_fit_equal(C.byref(a), C.byref(b))
    base = make_params(model=1, radius=0.3)
    for kw in (dict(tangent=1), dict(dofalloff=1), dict(falloffrate=3.0), dict(falloffradius=0.5), dict(maxedges=9),
               dict(morphspace=1), dict(doclampweight=1), dict(group="0-10"), dict(strict_reference=1), dict(layers=7),
               dict(qcoef=2.0)):
        assert eq(base, make_params(model=1, radius=0.3, **kw)) == 1, kw
    for kw in (dict(radius=0.31), dict(term=1), dict(kernel=1), dict(model=0), dict(eval_precision=2), dict(eval_path=1),
               dict(factor_precision=1), dict(fidelity=1), dict(eval_tolerance=1e-6), {"lambda": 0.5}):
        q = make_params(**{**dict(model=1, radius=0.3), **kw})
        assert eq(base, q) == 0, kw
    qnn = make_params(model=0)
    assert eq(qnn, make_params(model=0, radius=7.0)) == 1      # QNN: `radius` is only the capture / falloff radius
    assert eq(qnn, make_params(model=0, zcoef=4.0)) == 0
    v1 = make_params(model=1, fidelity=1, layers=3)
    assert eq(v1, make_params(model=1, fidelity=1, layers=4)) == 0
