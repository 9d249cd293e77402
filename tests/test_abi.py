"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/peritext_b200.h declares; without a GPU it fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "peritext_b200.h")).read()
    return sorted(set(re.findall(r"\b(pt_[a-z_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from peritext_b200 import engine
    lib = engine.load_library()
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(engine.EXPORTS) == syms
    assert b"sm_100a" in lib.pt_version()


def test_struct_layouts_match_header():
    from peritext_b200 import packing as p
    assert p.INSDEL_DT.itemsize == 16 and p.MARK_DT.itemsize == 32 and p.DESC_DT.itemsize == 32
    assert p.RESULT_DT.itemsize == 32 and p.SPAN_DT.itemsize == 16
    assert p.MARK_DT.fields["attr"][1] == 20 and p.MARK_DT.fields["arrival"][1] == 24


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from peritext_b200.engine import BatchEngine, EngineError
    with pytest.raises(EngineError, match="no CPU fallback"):
        BatchEngine(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under peritext_b200/ may reference it."""
    pkg = os.path.join(ROOT, "peritext_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), f
                assert "oracle/" not in txt.replace("tests/", ""), f
