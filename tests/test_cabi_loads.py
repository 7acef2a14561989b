"""CPU: the C-ABI library builds/loads and exports every symbol include/af2b200.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from alphafold2_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "af2b200.h")).read()
    names = set(re.findall(r"\b(af2_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/af2b200.h but not exported"
    from alphafold2_b200 import _lib
    assert names == set(_lib.EXPORTED_SYMBOLS)


def test_abi_version(lib):
    assert lib.af2_abi_version() == 2
    assert lib.af2_last_error() is not None


def test_workspace_queries(lib):
    assert lib.af2_feed_forward_workspace(1024, 256, 1024) > 1024 * 256 * 2
    assert lib.af2_axial_attention_workspace(1, 128, 256, 256, 8, 64, 1) > 0
    assert lib.af2_triangle_multiply_workspace(1, 256, 256) > 0
    assert lib.af2_outer_mean_workspace(1, 128, 256, 256) > 0


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "alphafold2_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{f} imports the oracle"


def test_state_dict_keys_match_reference_fixture():
    """The fixture's state dict comes from the unmodified reference: every key must load into the drop-in."""
    import torch
    import alphafold2_b200 as A
    from conftest import load_golden
    fx = load_golden("alphafold2_distogram")
    model = A.Alphafold2(**fx["cfg"])
    res = model.load_state_dict(fx["state"], strict=False)
    assert not res.unexpected_keys
    fx = load_golden("evoformer_block")
    c = fx["cfg"]
    blk = A.EvoformerBlock(dim=c["dim"], seq_len=c["N"], heads=c["heads"], dim_head=c["dim_head"], attn_dropout=0., ff_dropout=0.)
    blk.load_state_dict(fx["state"], strict=True)


def test_cpu_tensor_raises():
    import torch
    import alphafold2_b200 as A
    ff = A.FeedForward(dim=32)
    with pytest.raises(RuntimeError):
        ff(torch.randn(2, 32))
