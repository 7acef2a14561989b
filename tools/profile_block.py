"""One EvoformerBlock of the C2 workload inside a cudaProfiler range (for `ncu --profile-from-start off`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphafold2_b200 as A  # noqa: E402
from bench import CFG, N_RES, N_SEQ, randomize_zero_init_  # noqa: E402

N = int(os.environ.get("AF2_N", N_RES))
S = int(os.environ.get("AF2_S", N_SEQ))
torch.manual_seed(0)
blk = A.EvoformerBlock(dim=CFG["dim"], seq_len=N, heads=CFG["heads"], dim_head=CFG["dim_head"], attn_dropout=0., ff_dropout=0.)
randomize_zero_init_(blk)
blk = blk.cuda().eval()
x = torch.randn(1, N, N, CFG["dim"], device="cuda")
m = torch.randn(1, S, N, CFG["dim"], device="cuda")
mask = torch.ones(1, N, N, dtype=torch.bool, device="cuda")
msa_mask = torch.ones(1, S, N, dtype=torch.bool, device="cuda")
blk.update_(x.clone(), m.clone(), mask, msa_mask)
torch.cuda.synchronize()
xx, mm = x.clone(), m.clone()
torch.cuda.synchronize()
torch.cuda.profiler.start()
blk.update_(xx, mm, mask, msa_mask)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
