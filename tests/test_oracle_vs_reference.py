"""-m "not gpu": the oracle against the LIVE, unmodified reference (dev container only; skipped on the GPU box,
where /root/reference does not exist).  Seeds / shapes differ from the committed fixtures of tests/golden/."""
import pytest
import torch

from oracle import evoformer_oracle as O
from oracle.ref_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")


@pytest.mark.parametrize("N,S,d,H,dh,depth,masked", [(24, 5, 32, 2, 16, 2, True), (17, 3, 64, 4, 16, 1, False)])
def test_evoformer_matches_live_reference_fp64(N, S, d, H, dh, depth, masked):
    ref = load_reference()
    torch.manual_seed(11 + N)
    evo = ref.Evoformer(depth=depth, dim=d, seq_len=N, heads=H, dim_head=dh, attn_dropout=0., ff_dropout=0.).eval()
    st = O.randomize_zero_init_({k: v.clone() for k, v in evo.state_dict().items()}, seed=N)
    evo.load_state_dict(st)
    evo = evo.double()
    x, m = torch.randn(1, N, N, d).double(), torch.randn(1, S, N, d).double()
    mask = msa_mask = None
    if masked:
        m1 = torch.ones(1, N, dtype=torch.bool)
        m1[:, -3:] = False
        mask = m1[:, :, None] & m1[:, None, :]
        msa_mask = torch.rand(1, S, N) > 0.15
        msa_mask[:, :, -3:] = False
        msa_mask[:, 0, :N - 3] = True
    with torch.no_grad():
        rx, rm = evo(x, m, mask=mask, msa_mask=msa_mask)
    st64 = {k: v.double() for k, v in st.items()}
    ox, om = O.evoformer(st64, "", x, m, H, depth, mask, msa_mask)
    # fp64 vs fp64; Q3's count+eps is evaluated in fp32 by the reference -> ~1e-7 relative on the pair track
    assert (ox - rx).abs().max().item() <= 1e-6 * max(1.0, rx.abs().max().item())
    assert (om - rm).abs().max().item() <= 1e-6 * max(1.0, rm.abs().max().item())


def test_alphafold2_distogram_matches_live_reference_fp32():
    ref = load_reference()
    torch.manual_seed(5)
    cfg = dict(dim=32, depth=1, heads=2, dim_head=16)
    model = ref.Alphafold2(**cfg).eval()
    st = O.randomize_zero_init_({k: v.clone() for k, v in model.state_dict().items()}, seed=3)
    model.load_state_dict(st)
    seq = torch.randint(0, 21, (1, 20))
    msa = torch.randint(0, 21, (1, 3, 20))
    mask = torch.ones(1, 20, dtype=torch.bool)
    mask[:, -2:] = False
    msa_mask = torch.ones(1, 3, 20, dtype=torch.bool)
    msa_mask[:, :, -2:] = False
    with torch.no_grad():
        r = model(seq, msa, mask=mask, msa_mask=msa_mask).distance
    o = O.alphafold2_distogram(st, seq, msa, mask, msa_mask, heads=cfg["heads"], depth=cfg["depth"])
    assert (o - r).abs().max().item() <= 2e-4 * max(1.0, r.abs().max().item())
