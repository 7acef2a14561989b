"""Timeline of CTA 0 of one attention launch (AF2_ATTN_TRACE=1): per key block, when the MMA warp issued S and P V and when
softmax warp 2 acquired S / published P.   AF2_ATTN_TRACE=1 python tools/attn_trace.py [N] [rows] > profiles/..."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("AF2_ATTN_TRACE", "1")
import alphafold2_b200 as A  # noqa: E402
from alphafold2_b200 import _lib  # noqa: E402
from bench import randomize_zero_init_  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
ax = A.AxialAttention(dim=256, heads=8, dim_head=64, row_attn=True, col_attn=False, accept_edges=True)
randomize_zero_init_(ax)
ax = ax.cuda().eval()
x = torch.randn(1, N, N, 256, device="cuda")
for _ in range(2):
    ax.add_to_(x.clone(), x, None)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_longlong * 1024)()
_lib.check(lib.af2_debug_attn_trace(buf))          # clears
ax.add_to_(x.clone(), x, None)
_lib.check(lib.af2_debug_attn_trace(buf))
t = list(buf)
t0 = min(v for v in t if v > 0)
print(f"# N={N}: per key block g of CTA 0 (cycles since the first stamp): S_issue[begin end]  PV_issue[begin end]  S_acquired  P_published")
prev = None
for g in range(64):
    r = t[g * 16: g * 16 + 16]
    if not any(r[:6]):
        break
    rel = [(v - t0) if v > 0 else -1 for v in r]
    period = "" if prev is None or rel[5] < 0 else f"  period {rel[5] - prev}"
    prev = rel[5] if rel[5] >= 0 else prev
    print(f"g={g:3d}  S[{rel[0]:7d} {rel[1]:7d}]  PV[{rel[2]:7d} {rel[3]:7d}]  S_acq {rel[4]:7d}  P_pub {rel[5]:7d}  softmax {rel[5] - rel[4]:5d}  K_issue {rel[6]:7d}  MMA_free {rel[7]:7d} probe {rel[10]:7d} issue_s {rel[8]:7d} K_ok {rel[9]:7d}{period}")
