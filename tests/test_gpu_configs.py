"""-m gpu: parity AT THE CONFIGURATIONS THAT ARE BENCHMARKED (BASELINE.json configs): C2 = dim 256, depth 12, heads 8,
dim_head 64, N_res 256, MSA 128x256 through Alphafold2.forward and Evoformer.forward; one block at the C3 (N 384, MSA 512)
and C4 (N 512, MSA 1024) shapes.  Reference path: alphafold2.py:448-467 driven from :802-807.

The oracle (oracle/evoformer_oracle.py, pinned to the reference by tests/test_oracle_golden.py) is evaluated in fp64 ON THE
GPU here -- it is device-generic torch code and at these sizes a CPU evaluation takes minutes per block; a small case below
pins cuda-fp64 == cpu-fp64.  Masks are the partial masks of SURVEY.md 8(d): the last N/8 residues padded, ~10 % of the
MSA rows fully masked."""
import pytest
import torch

from gpu_util import autocast_yardstick, check, depth_gates, dev
from oracle import evoformer_oracle as O

pytestmark = pytest.mark.gpu

C2 = dict(dim=256, depth=12, heads=8, dim_head=64)


def _masks(B, N, S, seed):
    g = torch.Generator().manual_seed(seed)
    mask1 = torch.ones(B, N, dtype=torch.bool)
    mask1[:, -(N // 8):] = False
    msa_mask = torch.ones(B, S, N, dtype=torch.bool)
    dead = torch.rand(B, S, generator=g) < 0.1
    dead[:, 0] = False
    msa_mask[dead] = False
    msa_mask[:, :, -(N // 8):] = False
    return mask1, msa_mask


def _model(depth, seed=0):
    import alphafold2_b200 as A
    torch.manual_seed(seed)
    model = A.Alphafold2(**{**C2, "depth": depth})
    st = O.randomize_zero_init_({k: v.clone() for k, v in model.state_dict().items()})
    model.load_state_dict(st)
    return model.cuda().eval(), st


def test_oracle_cuda_fp64_equals_cpu_fp64():
    import alphafold2_b200 as A
    blk = A.EvoformerBlock(dim=64, seq_len=48, heads=2, dim_head=32, attn_dropout=0., ff_dropout=0.)
    st = O.randomize_zero_init_({k: v.clone() for k, v in blk.state_dict().items()})
    g = torch.Generator().manual_seed(1)
    x, m = torch.randn(1, 48, 48, 64, generator=g).double(), torch.randn(1, 6, 48, 64, generator=g).double()
    mask1, mm = _masks(1, 48, 6, 2)
    mask = mask1[:, :, None] & mask1[:, None, :]
    cx, cm = O.evoformer_block(dev(st, torch.float64, "cpu"), "", x, m, 2, mask, mm)
    gx, gm = O.evoformer_block(dev(st, torch.float64), "", x.cuda(), m.cuda(), 2, mask.cuda(), mm.cuda())
    assert (cx - gx.cpu()).abs().max().item() < 1e-10 and (cm - gm.cpu()).abs().max().item() < 1e-10


def test_c2_alphafold2_forward_depth12():
    """The bench's e2e call: Alphafold2(dim=256, depth=12, heads=8, dim_head=64).forward(seq, msa, mask, msa_mask)."""
    model, st = _model(12)
    N, S = 256, 128
    g = torch.Generator().manual_seed(7)
    seq = torch.randint(0, 21, (1, N), generator=g)
    msa = torch.randint(0, 21, (1, S, N), generator=g)
    mask1, msa_mask = _masks(1, N, S, 8)
    ret = model(seq.cuda(), msa.cuda(), mask=mask1.cuda(), msa_mask=msa_mask.cuda())
    args = (seq.cuda(), msa.cuda(), mask1.cuda(), msa_mask.cuda(), C2["heads"], 12)
    with torch.no_grad():
        ref = O.alphafold2_distogram(dev(st, torch.float64), *args, dtype=torch.float64, chunk=64)
        ac = autocast_yardstick(lambda: O.alphafold2_distogram(dev(st), *args, chunk=64))
    assert tuple(ret.distance.shape) == (1, N, N, 37)
    check("C2/alphafold2_forward/depth12/distance", ret.distance, ref, ac.float(), **depth_gates(12))


def test_c2_evoformer_depth12():
    """The bench's device-timed call: Evoformer.forward(x, m, mask, msa_mask) at C2, unit-normal inputs."""
    model, st = _model(12, seed=1)
    w = {k[len("net."):]: v for k, v in st.items() if k.startswith("net.")}
    N, S, d = 256, 128, 256
    g = torch.Generator().manual_seed(11)
    x, m = torch.randn(1, N, N, d, generator=g), torch.randn(1, S, N, d, generator=g)
    mask1, msa_mask = _masks(1, N, S, 12)
    mask = mask1[:, :, None] & mask1[:, None, :]
    xo, mo = model.net(x.cuda(), m.cuda(), mask=mask.cuda(), msa_mask=msa_mask.cuda())
    with torch.no_grad():
        rx, rm = O.evoformer(dev(w, torch.float64), "", x.double().cuda(), m.double().cuda(), 8, 12, mask.cuda(), msa_mask.cuda(), chunk=64)
        ax, am = autocast_yardstick(lambda: O.evoformer(dev(w), "", x.cuda(), m.cuda(), 8, 12, mask.cuda(), msa_mask.cuda(), chunk=64))
    check("C2/evoformer/depth12/x", xo, rx, ax.float(), **depth_gates(12))
    check("C2/evoformer/depth12/m", mo, rm, am.float(), **depth_gates(12))


@pytest.mark.parametrize("tag,N,S", [("C3", 384, 512), ("C4", 512, 1024)])
def test_block_at_config_shape(tag, N, S):
    """One EvoformerBlock at the C3 / C4 shapes (n > 256: streamed-bias attention, multi-wave contractions)."""
    import alphafold2_b200 as A
    d, H, dh = 256, 8, 64
    torch.manual_seed(3)
    blk = A.EvoformerBlock(dim=d, seq_len=N, heads=H, dim_head=dh, attn_dropout=0., ff_dropout=0.)
    st = O.randomize_zero_init_({k: v.clone() for k, v in blk.state_dict().items()})
    blk.load_state_dict(st)
    blk = blk.cuda().eval()
    g = torch.Generator().manual_seed(5)
    x, m = torch.randn(1, N, N, d, generator=g), torch.randn(1, S, N, d, generator=g)
    mask1, msa_mask = _masks(1, N, S, 6)
    mask = mask1[:, :, None] & mask1[:, None, :]
    xo, mo, _, _ = blk((x.cuda(), m.cuda(), mask.cuda(), msa_mask.cuda()))
    with torch.no_grad():
        rx, rm = O.evoformer_block(dev(st, torch.float64), "", x.double().cuda(), m.double().cuda(), H, mask.cuda(), msa_mask.cuda(), chunk=16)
        ax, am = autocast_yardstick(lambda: O.evoformer_block(dev(st), "", x.cuda(), m.cuda(), H, mask.cuda(), msa_mask.cuda(), chunk=16))
    check(f"{tag}/evoformer_block/x", xo, rx, ax.float())
    check(f"{tag}/evoformer_block/m", mo, rm, am.float())
