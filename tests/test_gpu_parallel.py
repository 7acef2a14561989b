"""-m gpu: the sharded schedule with the CUDA stage ops.  With one rank (NCCL group of size 1) it must agree with the
monolithic single-GPU modules and the fp64 oracle; tests/parallel_check_multi_gpu.py repeats this on 2+ GPUs under torchrun."""
import os

import pytest
import torch
import torch.distributed as dist

from gpu_util import check, to64
from oracle import evoformer_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pg():
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29541")
        torch.cuda.set_device(0)
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    yield None
    if dist.is_initialized():
        dist.destroy_process_group()


@pytest.mark.parametrize("d,H,dh,N,S,masked", [(64, 2, 32, 48, 8, True), (128, 4, 32, 64, 4, False), (256, 8, 64, 136, 40, True)])
def test_sharded_one_rank(pg, d, H, dh, N, S, masked):
    import alphafold2_b200 as A
    from alphafold2_b200.parallel import sharded_evoformer_forward
    torch.manual_seed(0)
    evo = A.Evoformer(depth=2, dim=d, seq_len=N, heads=H, dim_head=dh, attn_dropout=0., ff_dropout=0.)
    st = O.randomize_zero_init_({k: v.clone() for k, v in evo.state_dict().items()}, std=0.05)
    evo.load_state_dict(st)
    evo = evo.cuda().eval()
    x, m = torch.randn(1, N, N, d), torch.randn(1, S, N, d)
    mask = msa_mask = None
    if masked:
        m1 = torch.ones(1, N, dtype=torch.bool); m1[:, -N // 8:] = False
        mask = m1[:, :, None] & m1[:, None, :]
        msa_mask = torch.rand(1, S, N) > 0.1
        msa_mask[:, 0] = True
    cu = lambda t: None if t is None else t.cuda()  # noqa: E731
    xs, ms = sharded_evoformer_forward(evo, x.cuda(), m.cuda(), cu(mask), cu(msa_mask))
    x1, m1_ = evo(x.cuda(), m.cuda(), mask=cu(mask), msa_mask=cu(msa_mask))
    rx, rm = O.evoformer(to64(st), "", x.double(), m.double(), H, 2, mask, msa_mask)
    check(f"sharded_p1/d{d}N{N}/x", xs, rx)
    check(f"sharded_p1/d{d}N{N}/m", ms, rm)
    # same kernels, same order: the two paths may differ only through tile-shape dependent accumulation order
    assert (xs - x1).abs().max().item() <= 2e-3 * rx.abs().max().item()
    assert (ms - m1_).abs().max().item() <= 2e-3 * rm.abs().max().item()
