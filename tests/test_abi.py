"""CPU-side checks of the drop-in boundary: libfqsk.so loads, exports every symbol include/fqsk.h declares, and the
product path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

from fqsqueezer_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "fqsk.h")).read()
    # prototypes only (a line that starts with the return type): inline helpers of fqsk_ctx.h mentioned in comments are not exports
    return sorted(set(re.findall(r"^(?:int|void|const char \*)\s*(fqsk_[a-z0-9_]+)\s*\(", src, re.M)))


def test_library_exports_every_declared_symbol():
    from fqsqueezer_b200 import build
    build.build()
    lib = C.CDLL(E.LIB_PATH)
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(E.EXPORTS) == names


def test_struct_layouts_match_header(tmp_path):
    """sizeof() as the C compiler sees include/fqsk.h must equal the ctypes / numpy mirrors."""
    import subprocess
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "fqsk.h"\nint main(void){printf("%zu %zu %zu %zu\\n", sizeof(fqsk_params), sizeof(fqsk_stats), sizeof(fqsk_base_rec), sizeof(fqsk_read_desc));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert sizes == [C.sizeof(E._Params), C.sizeof(E._Stats), E.REC_DTYPE.itemsize, E.READ_DESC_DTYPE.itemsize]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(E.FqskError) as ei:
        E.KmerEngine(14, 17, 19, 9)
    assert ei.value.code == -2     # FQSK_E_NO_DEVICE


def test_product_package_does_not_import_oracle():
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle|from\s+\.+oracle)|libfqs_oracle|oracle/_ref", re.M)
    for fn in os.listdir(os.path.join(ROOT, "fqsqueezer_b200")):
        if fn.endswith(".py"):
            assert not pat.search(open(os.path.join(ROOT, "fqsqueezer_b200", fn)).read()), fn
    for fn in os.listdir(os.path.join(ROOT, "fqsqueezer_b200", "csrc")):
        assert "oracle" not in open(os.path.join(ROOT, "fqsqueezer_b200", "csrc", fn)).read().lower(), fn


def test_reference_side_binding_uses_only_the_declared_abi():
    """host/fqsk_live.h (the binding compiled into the reference) resolves its entry points by name: every one of them must be
    declared in include/fqsk.h and exported by libfqsk.so; the binding and its build recipe never reach for the oracle."""
    from fqsqueezer_b200 import build
    build.build()
    src = open(os.path.join(ROOT, "host", "fqsk_live.h")).read()
    bound = sorted(set(re.findall(r'sym\(p_[a-z_]+, "(fqsk_[a-z0-9_]+)"\)', src)))
    assert len(bound) >= 9
    lib = C.CDLL(E.LIB_PATH)
    declared = _declared()
    for n in bound:
        assert n in declared and hasattr(lib, n), n
    for fn in ("fqsk_live.h", "build_host.py"):
        assert not re.search(r"libfqs_oracle|fqso_|oracle/|oracle\.|import\s+oracle|from\s+oracle", open(os.path.join(ROOT, "host", fn)).read()), fn
