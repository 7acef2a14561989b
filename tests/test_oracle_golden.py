"""CPU: pin the oracle restatement (oracle/evoformer_oracle.py) to golden vectors produced by the
unmodified reference (oracle/make_golden.py).  fp64 vs fp64 must agree to ~1e-12; fp32 vs fp32 to
accumulation-order noise (~1e-6)."""
import pytest
import torch

from oracle import evoformer_oracle as O
from conftest import load_golden


def _to(w, dt):
    return {k: (v.to(dt) if torch.is_floating_point(v) else v) for k, v in w.items()}


def _check(out, fx, dt):
    ref = fx["out_fp64"] if dt == torch.float64 else fx["out_fp32"]
    tol = dict(rtol=1e-9, atol=1e-11) if dt == torch.float64 else dict(rtol=2e-4, atol=2e-5)
    if isinstance(ref, (tuple, list)):
        for o, r in zip(out, ref):
            torch.testing.assert_close(o, r.to(o.dtype), **tol)
    else:
        torch.testing.assert_close(out, ref.to(out.dtype), **tol)


DT = [torch.float64, torch.float32]


@pytest.mark.parametrize("dt", DT)
def test_feed_forward(dt):
    fx = load_golden("feed_forward")
    _check(O.feed_forward(_to(fx["state"], dt), "", fx["inputs"]["x"].to(dt)), fx, dt)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("name,row", [("axial_row_edges_masked", True), ("axial_col_masked", False),
                                      ("axial_col_edges_pair", False), ("axial_row_nomask", True)])
def test_axial_attention(name, row, dt):
    fx = load_golden(name)
    i = fx["inputs"]
    e = i.get("edges")
    out = O.axial_attention(_to(fx["state"], dt), "", i["x"].to(dt), fx["cfg"]["heads"], row,
                            None if e is None else e.to(dt), i.get("mask"))
    _check(out, fx, dt)


@pytest.mark.parametrize("dt", DT)
def test_axial_attention_chunked(dt):
    fx = load_golden("axial_row_edges_masked")
    i = fx["inputs"]
    out = O.axial_attention(_to(fx["state"], dt), "", i["x"].to(dt), fx["cfg"]["heads"], True,
                            i["edges"].to(dt), i["mask"], chunk=3)
    _check(out, fx, dt)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("mix", ["outgoing", "ingoing"])
def test_triangle_multiply(mix, dt):
    fx = load_golden(f"triangle_multiply_{mix}")
    i = fx["inputs"]
    _check(O.triangle_multiply(_to(fx["state"], dt), "", i["x"].to(dt), mix, i["mask"]), fx, dt)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("literal", [False, True])
@pytest.mark.parametrize("name", ["outer_mean_masked", "outer_mean_nomask"])
def test_outer_mean(name, literal, dt):
    fx = load_golden(name)
    i = fx["inputs"]
    out = O.outer_mean(_to(fx["state"], dt), "", i["m"].to(dt), i.get("mask"), literal=literal)
    ref = fx["out_fp64"] if dt == torch.float64 else fx["out_fp32"]
    # Q3: count+eps is evaluated in fp32 by the reference even in an fp64 run -> 1e-7 relative
    tol = dict(rtol=1e-6, atol=1e-8) if dt == torch.float64 else dict(rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(out, ref.to(out.dtype), **tol)


@pytest.mark.parametrize("dt", DT)
def test_evoformer_block(dt):
    fx = load_golden("evoformer_block")
    i = fx["inputs"]
    x, m = O.evoformer_block(_to(fx["state"], dt), "", i["x"].to(dt), i["m"].to(dt),
                             fx["cfg"]["heads"], i["mask"], i["msa_mask"])
    ref = fx["out_fp64"] if dt == torch.float64 else fx["out_fp32"]
    tol = dict(rtol=1e-6, atol=1e-7) if dt == torch.float64 else dict(rtol=5e-4, atol=5e-5)
    torch.testing.assert_close(x, ref[0].to(dt), **tol)
    torch.testing.assert_close(m, ref[1].to(dt), **tol)


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("name,depth", [("evoformer_depth2", 2), ("evoformer_nomask", 1)])
def test_evoformer(name, depth, dt):
    fx = load_golden(name)
    i = fx["inputs"]
    x, m = O.evoformer(_to(fx["state"], dt), "", i["x"].to(dt), i["m"].to(dt), fx["cfg"]["heads"],
                       depth, i.get("mask"), i.get("msa_mask"))
    ref = fx["out_fp64"] if dt == torch.float64 else fx["out_fp32"]
    tol = dict(rtol=1e-6, atol=1e-7) if dt == torch.float64 else dict(rtol=5e-4, atol=5e-5)
    torch.testing.assert_close(x, ref[0].to(dt), **tol)
    torch.testing.assert_close(m, ref[1].to(dt), **tol)


def test_alphafold2_distogram():
    fx = load_golden("alphafold2_distogram")
    i, cfg = fx["inputs"], fx["cfg"]
    out = O.alphafold2_distogram(fx["state"], i["seq"], i["msa"], i["mask"], i["msa_mask"],
                                 cfg["heads"], cfg["depth"])
    torch.testing.assert_close(out, fx["out_fp32"], rtol=5e-4, atol=5e-5)
    out = O.alphafold2_distogram(_to(fx["state"], torch.float64), i["seq"], i["msa"], i["mask"],
                                 i["msa_mask"], cfg["heads"], cfg["depth"], dtype=torch.float64)
    torch.testing.assert_close(out, fx["out_fp64"], rtol=1e-6, atol=1e-7)
    out = O.alphafold2_distogram(fx["state"], i["seq"], None, i["mask"], None, cfg["heads"], cfg["depth"])
    torch.testing.assert_close(out, fx["out_fp32_no_msa"], rtol=5e-4, atol=5e-5)


def test_rotary():
    fx = load_golden("rotary")
    i = fx["inputs"]
    torch.testing.assert_close(O.apply_rotary_pos_emb(i["x"], i["sin"], i["cos"]), fx["out_fp32"], rtol=0, atol=0)
    s, c = O.fixed_positional_embedding(24, 10)
    torch.testing.assert_close(s, i["sin"]); torch.testing.assert_close(c, i["cos"])


def test_quirk_masked_query_uniform():
    """Q1: a masked query row attends uniformly over ALL keys (masked ones included)."""
    torch.manual_seed(0)
    d, H, dh, n = 16, 2, 8, 6
    w = {"to_q.weight": torch.randn(H * dh, d), "to_kv.weight": torch.randn(2 * H * dh, d),
         "gating.weight": torch.zeros(H * dh, d), "gating.bias": torch.full((H * dh,), 50.),
         "to_out.weight": torch.eye(d, H * dh), "to_out.bias": torch.zeros(d)}
    x = torch.randn(1, n, d)
    mask = torch.tensor([[True, True, False, True, False, True]])
    out = O.attention(w, "", x, H, mask)
    v = (x @ w["to_kv.weight"].T)[..., H * dh:]
    torch.testing.assert_close(out[0, 2], v[0].mean(0), rtol=1e-5, atol=1e-6)


def test_flops_formula():
    assert abs(O.evoformer_flops_per_block(256, 128, 256, 8, 64) / 1e9 - 649.3) < 0.5


def test_evoformer_global_column_attn():
    """n2: tied-query ("global") ingoing triangle attention of the extra-MSA stack (alphafold2.py:142-151, 250, 367)."""
    fx = load_golden("evoformer_global_col")
    i, c = fx["inputs"], fx["cfg"]
    x, m = O.evoformer(fx["state"], "", i["x"], i["m"], c["heads"], 1, i["mask"], i["msa_mask"], global_column_attn=True)
    torch.testing.assert_close(x, fx["out_fp32"][0], rtol=5e-4, atol=5e-5)
    torch.testing.assert_close(m, fx["out_fp32"][1], rtol=5e-4, atol=5e-5)
    x, m = O.evoformer(_to(fx["state"], torch.float64), "", i["x"].double(), i["m"].double(), c["heads"], 1, i["mask"], i["msa_mask"],
                       global_column_attn=True)
    torch.testing.assert_close(x, fx["out_fp64"][0], rtol=1e-6, atol=1e-7)
    # the tie really matters: the untied block gives a different pair tensor
    xu, _ = O.evoformer(fx["state"], "", i["x"], i["m"], c["heads"], 1, i["mask"], i["msa_mask"])
    assert (xu - fx["out_fp32"][0]).abs().max() > 1e-3


def test_alphafold2_extra_msa():
    """n2: Alphafold2.forward with extra_msa / extra_msa_mask (alphafold2.py:789-798, quirk Q11 included)."""
    fx = load_golden("alphafold2_extra_msa")
    i, cfg = fx["inputs"], fx["cfg"]
    out = O.alphafold2_distogram(fx["state"], i["seq"], i["msa"], i["mask"], i["msa_mask"], cfg["heads"], cfg["depth"],
                                 extra_msa_mask=i["extra_msa_mask"], extra_depth=cfg["extra_msa_evoformer_layers"])
    torch.testing.assert_close(out, fx["out_fp32"], rtol=5e-4, atol=5e-5)
    out = O.alphafold2_distogram(_to(fx["state"], torch.float64), i["seq"], i["msa"], i["mask"], i["msa_mask"], cfg["heads"],
                                 cfg["depth"], dtype=torch.float64, extra_msa_mask=i["extra_msa_mask"],
                                 extra_depth=cfg["extra_msa_evoformer_layers"])
    torch.testing.assert_close(out, fx["out_fp64"], rtol=1e-6, atol=1e-7)
