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


def test_run_compression_is_lossless():
    """pt_compress_runs (host code of the C-ABI library): expanding the runs gives back the records bit for bit."""
    import numpy as np
    from peritext_b200 import workload
    from peritext_b200.engine import compress_runs
    for cfg in ("c2", "c3", "c4"):
        b = workload.generate(cfg, n_docs=6, ops_per_doc=1500)
        r = compress_runs(b)
        out = np.zeros(len(b.insdel), b.insdel.dtype)
        for li in range(b.n_logs):
            o, t, k = int(b.desc[li]["insdel_off"]), int(r.tok_off[li]), 0
            for q in r.runs[int(r.run_off[li]): int(r.run_off[li + 1])]:
                cnt, kind = int(q["kind_count"]) & 0x3FFFFFFF, int(q["kind_count"]) >> 30
                for j in range(cnt):
                    if kind == 0:
                        out[o + k] = (int(q["ctr0"]) + j, int(q["ref_ctr"]) if j == 0 else int(q["ctr0"]) + j - 1, int(q["actor"]),
                                      int(q["ref_actor"]) if j == 0 else int(q["actor"]), int(r.tokens[t])); t += 1
                    else:
                        out[o + k] = (int(q["ctr0"]) + j, int(q["ref_ctr"]) + j, int(q["actor"]), int(q["ref_actor"]), 1 << 30)
                    k += 1
            assert k == int(b.desc[li]["n_insdel"])
        assert out.tobytes() == b.insdel.tobytes()
        assert r.nbytes < b.insdel.nbytes + b.marks.nbytes + b.desc.nbytes * 2
