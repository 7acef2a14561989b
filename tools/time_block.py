"""Device time of one EvoformerBlock at a BASELINE shape (CUDA events, L2 flushed between iterations):
   AF2_N=384 AF2_S=512 python tools/time_block.py  -> one JSON line {N, S, ms_per_block, tflops}"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphafold2_b200 as A  # noqa: E402
from bench import CFG, flops_per_block, randomize_zero_init_  # noqa: E402

N = int(os.environ.get("AF2_N", 256))
S = int(os.environ.get("AF2_S", 128))
it = int(os.environ.get("AF2_ITERS", 10))
torch.manual_seed(0)
blk = A.EvoformerBlock(dim=CFG["dim"], seq_len=N, heads=CFG["heads"], dim_head=CFG["dim_head"], attn_dropout=0., ff_dropout=0.)
randomize_zero_init_(blk)
blk = blk.cuda().eval()
if os.environ.get("AF2_PRECISION_BLOCK"):
    A.set_precision(blk, os.environ["AF2_PRECISION_BLOCK"])
x = torch.randn(1, N, N, CFG["dim"], device="cuda")
m = torch.randn(1, S, N, CFG["dim"], device="cuda")
mask = torch.ones(1, N, N, dtype=torch.bool, device="cuda")
msa_mask = torch.ones(1, S, N, dtype=torch.bool, device="cuda")
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    blk.update_(x.clone(), m.clone(), mask, msa_mask)
tot = 0.0
for _ in range(it):
    xx, mm = x.clone(), m.clone()
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    blk.update_(xx, mm, mask, msa_mask)
    b.record()
    b.synchronize()
    tot += a.elapsed_time(b)
ms = tot / it
fl = flops_per_block(N, S, CFG["dim"], CFG["heads"], CFG["dim_head"])
print(json.dumps({"N": N, "S": S, "ms_per_block": ms, "tflops": fl / (ms * 1e-3) / 1e12,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("AF2_")}}))
