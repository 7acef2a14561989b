"""Timeline of one CTA of the fused projection kernel (debug aid; run on the GPU box with AF2_PROJ_TRACE=1):
     AF2_PROJ_TRACE=1 python tools/proj_trace.py [ff|attn|tri]
   Prints, per accumulator tile of cluster 0's leader CTA, when the MMA warp got the accumulator stage, when the first
   weight stage had landed, when the last MMA of the tile was issued, and when epilogue warps 4 / 8 started waiting for the
   tile, got it and finished it; per item, when producer warp 12 got the A buffer and when it finished.  Cycles are
   relative to the kernel start (clock64 of that SM)."""
import ctypes as C
import os
import sys
import torch

os.environ.setdefault("AF2_PROJ_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alphafold2_b200.alphafold2 as A
from alphafold2_b200 import _lib

which = sys.argv[1] if len(sys.argv) > 1 else "ff"
torch.manual_seed(0)
dev = "cuda:0"
N, d = 256, 256
x = torch.randn(1, N, N, d, device=dev)
with torch.no_grad():
    if which == "ff":
        mod = A.FeedForward(dim=d).to(dev)
    elif which == "attn":
        mod = A.AxialAttention(dim=d, heads=8, dim_head=64, row_attn=True, col_attn=False, accept_edges=True).to(dev)
    else:
        mod = A.TriangleMultiplicativeModule(dim=d, mix="outgoing").to(dev)
    for p in mod.parameters():
        if p.dim() > 0 and float(p.abs().sum()) == 0.0:
            p.normal_(0, 0.02)
    args = (x,) if which != "attn" else (x,)
    kw = {"edges": x} if which == "attn" else {}
    for _ in range(3):
        y = mod(*args, **kw)
    torch.cuda.synchronize()
buf = (C.c_longlong * 2048)()
_lib.check(_lib.load().af2_debug_proj_trace(buf))
t = list(buf)
t0 = t[1023]
rel = lambda v: (v - t0) if v else None
print(f"kind={which}  (cycles since kernel start; leader CTA of cluster 0)")
print(" tile | mma: acc_free  w_landed  issued | epi4: wait   got    done | epi8: wait   got    done")
for k in range(40):
    m = [rel(t[3 * k + i]) for i in range(3)]
    e4 = [rel(t[256 + 3 * k + i]) for i in range(3)]
    e8 = [rel(t[512 + 3 * k + i]) for i in range(3)]
    if m[0] is None and e4[0] is None:
        break
    f = lambda a: " ".join(f"{v:8d}" if v is not None else "       -" for v in a)
    print(f" {k:4d} | {f(m)} | {f(e4)} | {f(e8)}")
print(" item | producer12: got_buffer  done")
for it in range(16):
    a, b = rel(t[768 + 2 * it]), rel(t[768 + 2 * it + 1])
    if a is None and b is None:
        break
    print(f" {it:4d} | {a}  {b}")
print(" chunk (epilogue warp 4) | loads landed -> packed -> store buffer free -> staged + fenced -> TMA store issued   (deltas)")
prev = None
for c in range(60):
    v = [t[1024 + 5 * c + i] for i in range(5)]
    if not v[0]:
        break
    d = [v[i + 1] - v[i] for i in range(4)]
    gap = (v[0] - prev) if prev else 0
    prev = v[4]
    print(f" {c:4d} | t={v[0] - t0:8d}  gap_from_prev={gap:6d}  pack={d[0]:5d} wait_buf={d[1]:5d} sts_fence={d[2]:5d} issue={d[3]:5d}")
