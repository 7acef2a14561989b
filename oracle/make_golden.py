"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.pt from the UNMODIFIED reference.

Run in the dev container (needs /root/reference):   python -m oracle.make_golden
The reference's own tests hold no numeric vectors, so the golden vectors are
outputs of the live reference modules on seeded inputs, in fp32 (parity target) and fp64
(ground truth), with the Q8 randomisation applied so that no path is an identity.

Each fixture is a dict: {"cfg": {...}, "state": state_dict (fp32), "inputs": {...},
"out_fp32": ..., "out_fp64": ..., "out_autocast_bf16": ... (where cheap)}.
Shapes are deliberately ragged (N not a multiple of 8, S tiny) and masks are partial.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference, load_reference_rotary  # noqa: E402
from oracle.evoformer_oracle import randomize_zero_init_  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _masks(b, S, N, seed):
    g = torch.Generator().manual_seed(seed)
    mask = torch.ones(b, N, dtype=torch.bool)
    mask[:, -max(1, N // 8):] = False                      # trailing residue padding
    msa_mask = torch.ones(b, S, N, dtype=torch.bool)
    msa_mask[:, :, -max(1, N // 8):] = False
    if S > 2:
        msa_mask[:, S - 1, :] = False                      # one fully masked MSA row
    msa_mask &= torch.rand(b, S, N, generator=g) > 0.1     # random holes
    msa_mask[:, 0, : N - max(1, N // 8)] = True
    return mask, msa_mask


def _run(mod, dtype, fn):
    mod = mod.to(dtype)
    with torch.no_grad():
        out = fn(mod, dtype)
    mod.to(torch.float32)
    return out


def _autocast(mod, fn):
    with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
        return fn(mod, torch.float32)


def main():
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)

    # ---------------- trunk-level fixtures: every module on the hot path -----------------
    d, H, dh, N, S, b = 64, 2, 32, 20, 5, 2
    cfg = dict(dim=d, heads=H, dim_head=dh, N=N, S=S, b=b)
    x = torch.randn(b, N, N, d)
    m = torch.randn(b, S, N, d)
    mask1, msa_mask = _masks(b, S, N, 7)
    mask = mask1[:, :, None] & mask1[:, None, :]

    def save(name, mod, inputs, fn, with_autocast=True):
        state = randomize_zero_init_({k: v.clone() for k, v in mod.state_dict().items()})
        mod.load_state_dict(state)
        mod.eval()
        fx = {"cfg": cfg, "state": state, "inputs": inputs}
        fx["out_fp32"] = _run(mod, torch.float32, fn)
        fx["out_fp64"] = _run(mod, torch.float64, fn)
        if with_autocast:
            fx["out_autocast_bf16"] = _autocast(mod, fn)
        torch.save(fx, os.path.join(OUT, name + ".pt"))
        print("wrote", name)

    c = lambda t, dt: t.to(dt)  # noqa: E731

    save("feed_forward", ref.FeedForward(dim=d), {"x": x},
         lambda mod, dt: mod(c(x, dt)))
    save("axial_row_edges_masked", ref.AxialAttention(dim=d, heads=H, dim_head=dh, row_attn=True, col_attn=False, accept_edges=True),
         {"x": m, "edges": x, "mask": msa_mask},
         lambda mod, dt: mod(c(m, dt), edges=c(x, dt), mask=msa_mask))
    save("axial_col_masked", ref.AxialAttention(dim=d, heads=H, dim_head=dh, row_attn=False, col_attn=True),
         {"x": m, "mask": msa_mask},
         lambda mod, dt: mod(c(m, dt), mask=msa_mask))
    save("axial_col_edges_pair", ref.AxialAttention(dim=d, heads=H, dim_head=dh, row_attn=False, col_attn=True, accept_edges=True),
         {"x": x, "edges": x, "mask": mask},
         lambda mod, dt: mod(c(x, dt), edges=c(x, dt), mask=mask))
    save("axial_row_nomask", ref.AxialAttention(dim=d, heads=H, dim_head=dh, row_attn=True, col_attn=False, accept_edges=True),
         {"x": x, "edges": x},
         lambda mod, dt: mod(c(x, dt), edges=c(x, dt)))
    for mix in ("outgoing", "ingoing"):
        save(f"triangle_multiply_{mix}", ref.TriangleMultiplicativeModule(dim=d, mix=mix),
             {"x": x, "mask": mask},
             lambda mod, dt: mod(c(x, dt), mask=mask))
    save("outer_mean_masked", ref.OuterMean(d), {"m": m, "mask": msa_mask},
         lambda mod, dt: mod(c(m, dt), mask=msa_mask))
    save("outer_mean_nomask", ref.OuterMean(d), {"m": m},
         lambda mod, dt: mod(c(m, dt)))
    save("evoformer_block", ref.EvoformerBlock(dim=d, seq_len=N, heads=H, dim_head=dh, attn_dropout=0., ff_dropout=0.),
         {"x": x, "m": m, "mask": mask, "msa_mask": msa_mask},
         lambda mod, dt: mod((c(x, dt), c(m, dt), mask, msa_mask))[:2])
    save("evoformer_depth2", ref.Evoformer(depth=2, dim=d, seq_len=N, heads=H, dim_head=dh, attn_dropout=0., ff_dropout=0.),
         {"x": x, "m": m, "mask": mask, "msa_mask": msa_mask},
         lambda mod, dt: mod(c(x, dt), c(m, dt), mask=mask, msa_mask=msa_mask))
    save("evoformer_nomask", ref.Evoformer(depth=1, dim=d, seq_len=N, heads=H, dim_head=dh, attn_dropout=0., ff_dropout=0.),
         {"x": x, "m": m},
         lambda mod, dt: mod(c(x, dt), c(m, dt)))

    # ---------------- model-API fixture: Alphafold2.forward -> distogram (tests/test_attention.py:8-27) ----
    torch.manual_seed(1)
    mcfg = dict(dim=32, depth=2, heads=2, dim_head=32)
    model = ref.Alphafold2(**mcfg)
    model.eval()
    full = randomize_zero_init_({k: v.clone() for k, v in model.state_dict().items()})
    model.load_state_dict(full)
    used = ("token_emb.", "to_pairwise_repr.", "pos_emb.", "net.", "to_distogram_logits.")
    state = {k: v for k, v in full.items() if k.startswith(used)}
    bb, n, s = 2, 24, 3
    seq = torch.randint(0, 21, (bb, n))
    msa = torch.randint(0, 21, (bb, s, n))
    mk, mmk = _masks(bb, s, n, 11)
    fx = {"cfg": mcfg, "state": state, "inputs": {"seq": seq, "msa": msa, "mask": mk, "msa_mask": mmk}}
    with torch.no_grad():
        fx["out_fp32"] = model(seq, msa, mask=mk, msa_mask=mmk).distance
        fx["out_fp32_no_msa"] = model(seq, mask=mk).distance
        model.double()
        fx["out_fp64"] = model(seq, msa, mask=mk, msa_mask=mmk).distance
        model.float()
    torch.save(fx, os.path.join(OUT, "alphafold2_distogram.pt"))
    print("wrote alphafold2_distogram")

    # ---------------- rotary.py (dead at HEAD; loaded standalone) -----------------------------------------
    rot = load_reference_rotary()
    torch.manual_seed(2)
    xq = torch.randn(2, 3, 10, 32)
    sin, cos = rot.FixedPositionalEmbedding(24)(10, xq.device)          # rot_dim 24 < dh 32: tail passes through
    fx = {"inputs": {"x": xq, "sin": sin, "cos": cos},
          "out_fp32": rot.apply_rotary_pos_emb(xq, (sin, cos)),
          "out_fp64": rot.apply_rotary_pos_emb(xq.double(), (sin.double(), cos.double()))}
    torch.save(fx, os.path.join(OUT, "rotary.pt"))
    print("wrote rotary")


if __name__ == "__main__":
    main()
