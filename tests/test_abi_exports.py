"""The C-ABI library loads on a CPU-only box and exports every symbol include/cffm_b200.h declares
(no compute call is made here: there is no GPU).  Also: the product package never imports oracle/."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from vss_cffm_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_is_built():
    assert os.path.exists(_abi.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'`"


def test_every_header_symbol_is_exported_and_bound():
    names = _abi.header_symbols()
    assert len(names) >= 18
    lib = ctypes.CDLL(_abi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_abi.SIGNATURES) == names, "ctypes signature table and header disagree"


def test_ctypes_arity_matches_header():
    with open(_abi.HEADER_PATH) as f:
        src = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    for name, (args, _) in _abi.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^)]*)\)", src)
        assert m, name
        params = [p.strip() for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))
        for p, a in zip(params, args):
            if "*" in p:
                assert a is ctypes.c_void_p, (name, p)
            elif p.startswith("int64_t"):
                assert a is ctypes.c_int64, (name, p)
            elif p.startswith("float"):
                assert a is ctypes.c_float, (name, p)
            elif p.startswith("int"):
                assert a is ctypes.c_int, (name, p)


def test_abi_version_and_error_text_without_gpu():
    lib = _abi.load()
    assert lib.cffm_abi_version() == 1
    import torch
    if not torch.cuda.is_available():
        assert lib.cffm_device_check() != 0          # fails loudly: no CPU fallback
        assert len(lib.cffm_last_error()) > 0
        with pytest.raises(_abi.CffmError):
            _abi.require_device()


def test_sass_contains_tcgen05_and_tma():
    """The GEMM really is a Blackwell-native kernel: UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA)."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _abi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vss_cffm_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
    code = "import sys; import vss_cffm_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT, timeout=300)


def test_cpu_tensors_are_rejected_not_silently_computed():
    import torch
    from vss_cffm_b200 import ops
    a = torch.zeros(8, 8, dtype=torch.float16)
    with pytest.raises(_abi.CffmError, match="no CPU fallback"):
        ops.gemm(a, a, out32=torch.zeros(8, 8))


def test_header_is_plain_c99(tmp_path):
    """include/cffm_b200.h is the drop-in boundary: it must compile as C (no C++ or torch types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "use.c"
    src.write_text('#include "cffm_b200.h"\nint main(void) { cffm_patch_or_version(); return 0; }\n'
                   .replace("cffm_patch_or_version()", "(void)cffm_abi_version"))
    r = subprocess.run([gcc, "-std=c99", "-pedantic", "-Werror", "-fsyntax-only", "-I", os.path.join(root, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
