"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads and exports every symbol that
include/scot_b200.h declares; the engine's parameter table equals the reference's state_dict layout."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def lib():
    from poseidon_b200 import _lib

    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "scot_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(scot_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 15
    for n in sorted(names):
        assert hasattr(lib, n), f"{n} declared in scot_b200.h but not exported by libscot_b200.so"
    assert lib.scot_abi_version() == 2


def test_errors_are_reported_not_thrown(lib):
    from poseidon_b200 import _lib

    d = _lib.ScotModelDesc()
    d.num_stages = 9  # unsupported
    h = ctypes.c_void_p()
    rc = lib.scot_engine_create(ctypes.byref(d), 1, ctypes.byref(h))
    assert rc == 1 and b"stages" in lib.scot_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "scot_engine_create")


@pytest.mark.parametrize("name", ["tiny", "tiny_ln", "T128", "B128"])
def test_engine_param_table_equals_reference_state_dict(name):
    """names/shapes come from fixtures recorded from the UNMODIFIED reference's state_dict()"""
    from poseidon_b200 import _lib
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    rec = torch.load(os.path.join(GOLD, f"{name}.pt"), weights_only=False)
    with torch.device("meta"):
        model = ScOT(ScOTConfig(**rec["config"]))
    eng = _lib.Engine(model._desc(), batch=2)
    assert set(eng.table) == set(rec["shapes"])
    spans = []
    for k, (off, numel, shape) in eng.table.items():
        assert tuple(shape) == tuple(rec["shapes"][k]), k
        assert off % 4 == 0 and off + numel <= eng.param_elems
        spans.append((off, off + numel))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])), "parameters overlap in the flat buffer"
    # q/k/v weights are contiguous (one [3C, C] GEMM operand), [bq | 0 | bv] likewise
    pre = "encoder.layers.0.blocks.0.attention.self."
    C = rec["shapes"][pre + "query.weight"][0]
    q, k, v = (eng.table[pre + n + ".weight"][0] for n in ("query", "key", "value"))
    assert k == q + C * C and v == k + C * C
    assert eng.table[pre + "value.bias"][0] == eng.table[pre + "query.bias"][0] + 2 * C
    assert eng.workspace_bytes > 0


def test_module_tree_matches_reference_keys_and_hf_roundtrip(tmp_path):
    from poseidon_b200.scOT.model import ConditionalLayerNorm, LayerNorm, ScOT, ScOTConfig

    rec = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
    m = ScOT(ScOTConfig(**rec["config"]))
    sd = m.state_dict()
    assert set(sd) == set(rec["shapes"]) and all(tuple(sd[k].shape) == tuple(rec["shapes"][k]) for k in sd)
    assert any(isinstance(x, ConditionalLayerNorm) for x in m.modules())
    m.save_pretrained(tmp_path)
    m2 = ScOT.from_pretrained(tmp_path)
    assert all(torch.equal(a, b) for a, b in zip(sd.values(), m2.state_dict().values()))
    rec2 = torch.load(os.path.join(GOLD, "tiny_ln.pt"), weights_only=False)
    m3 = ScOT(ScOTConfig(**rec2["config"]))
    assert any(isinstance(x, LayerNorm) for x in m3.modules()) and set(m3.state_dict()) == set(rec2["shapes"])


def test_no_cpu_fallback():
    from poseidon_b200.scOT.model import ScOT, ScOTConfig

    rec = torch.load(os.path.join(GOLD, "tiny.pt"), weights_only=False)
    m = ScOT(ScOTConfig(**rec["config"]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(pixel_values=torch.zeros(1, 3, 64, 64), time=torch.zeros(1))
    with pytest.raises(NotImplementedError):
        ScOT(ScOTConfig(**{**rec["config"], "drop_path_rate": 0.1}))


def test_product_path_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "poseidon_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle." not in src, f"{f} uses the oracle"
