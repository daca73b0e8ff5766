"""The C-ABI shared library loads and exports every symbol include/hmcb.h declares.
No compute calls: this runs on a box without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from hmclab_b200 import _build, _engine

    path = _build.build()  # no-op when the in-tree library is current
    return _engine.load_library(path)


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "hmcb.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hmcb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from hmclab_b200 import _engine

    declared = _declared_symbols()
    assert len(declared) >= 28
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in hmcb.h but not exported"
    assert set(declared) == set(_engine.SIGNATURES), "ctypes prototypes out of sync with hmcb.h"


def test_abi_version_and_error_channel(lib):
    assert lib.hmcb_abi_version() == 2
    assert isinstance(lib.hmcb_last_error(), bytes)


def test_block_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of hmcb_block as a C compiler sees the header == the ctypes mirror."""
    import shutil
    import subprocess

    from hmclab_b200._engine import _Block

    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    fields = [name for name, _ in _Block._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "hmcb.h"', "int main(void) {",
             '  printf("%zu ", sizeof(hmcb_block));']
    lines += [f'  printf("%zu ", offsetof(hmcb_block, {f}));' for f in fields]
    lines += ["  return 0;", "}"]
    src.write_text(chr(10).join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    numbers = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert numbers[0] == ctypes.sizeof(_Block)
    assert numbers[1:] == [getattr(_Block, f).offset for f in fields]


def test_create_fails_loudly_without_a_b200(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    handle = ctypes.c_void_p()
    status = lib.hmcb_create(0, 4, 3, ctypes.byref(handle))
    assert status != 0 and handle.value is None
    assert len(lib.hmcb_last_error()) > 0


def test_null_arguments_are_rejected_not_dereferenced(lib):
    assert lib.hmcb_set_integrator(None, 0, 10) != 0
    assert lib.hmcb_finalize(None) != 0
    assert lib.hmcb_misfit(None, None, None, None) != 0
    assert lib.hmcb_run_block(None, None, None) != 0
    assert lib.hmcb_destroy(None) == 0
    assert lib.hmcb_path(None) == -1


def test_product_has_no_cpu_fallback():
    """Without a CUDA device the Python engine refuses to run (and never imports oracle/)."""
    import sys

    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from hmclab_b200._engine import Engine, HmcbError

    plan = {"dims": 3, "terms": [], "checks": [], "likelihood": None, "reflect_lb": None,
            "reflect_ub": None}
    with pytest.raises(HmcbError, match="no CPU fallback"):
        Engine(plan, {"kind": "unit", "dims": 3}, 2)
    for mod in list(sys.modules):
        if mod.startswith("hmclab_b200"):
            src = getattr(sys.modules[mod], "__file__", "") or ""
            if src.endswith(".py"):
                with open(src) as f:
                    text = f.read()
                assert "import oracle" not in text and "from oracle" not in text, mod
