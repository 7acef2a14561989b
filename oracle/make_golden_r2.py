"""TEST INFRASTRUCTURE ONLY — round-2 fixtures from the UNMODIFIED reference (run in the dev container):
    python -m oracle.make_golden_r2
  tests/golden/evoformer_global_col.pt   Evoformer(depth 1, global_column_attn=True): tied-query ingoing triangle attention
                                         (alphafold2.py:142-151, 250, 367)
  tests/golden/alphafold2_extra_msa.pt   Alphafold2.forward(seq, msa, mask, msa_mask, extra_msa, extra_msa_mask)
                                         (alphafold2.py:789-798, including quirk Q11)
Same conventions as oracle/make_golden.py."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference  # noqa: E402
from oracle.evoformer_oracle import randomize_zero_init_  # noqa: E402
from oracle.make_golden import _masks, _run, _autocast, OUT  # noqa: E402


def main():
    ref = load_reference()
    torch.manual_seed(3)
    d, H, dh, N, S, b = 64, 2, 32, 20, 5, 2
    cfg = dict(dim=d, heads=H, dim_head=dh, N=N, S=S, b=b)
    x, m = torch.randn(b, N, N, d), torch.randn(b, S, N, d)
    mask1, msa_mask = _masks(b, S, N, 9)
    mask = mask1[:, :, None] & mask1[:, None, :]
    mod = ref.Evoformer(depth=1, dim=d, seq_len=N, heads=H, dim_head=dh, attn_dropout=0., ff_dropout=0., global_column_attn=True)
    state = randomize_zero_init_({k: v.clone() for k, v in mod.state_dict().items()})
    mod.load_state_dict(state)
    mod.eval()
    fn = lambda md, dt: md(x.to(dt), m.to(dt), mask=mask, msa_mask=msa_mask)  # noqa: E731
    fx = {"cfg": cfg, "state": state, "inputs": {"x": x, "m": m, "mask": mask, "msa_mask": msa_mask},
          "out_fp32": _run(mod, torch.float32, fn), "out_fp64": _run(mod, torch.float64, fn), "out_autocast_bf16": _autocast(mod, fn)}
    torch.save(fx, os.path.join(OUT, "evoformer_global_col.pt"))
    print("wrote evoformer_global_col")

    torch.manual_seed(4)
    mcfg = dict(dim=32, depth=1, heads=2, dim_head=32, extra_msa_evoformer_layers=2)
    model = ref.Alphafold2(**mcfg)
    model.eval()
    full = randomize_zero_init_({k: v.clone() for k, v in model.state_dict().items()})
    model.load_state_dict(full)
    used = ("token_emb.", "to_pairwise_repr.", "pos_emb.", "net.", "to_distogram_logits.", "extra_msa_evoformer.")
    state = {k: v for k, v in full.items() if k.startswith(used)}
    bb, n, s = 2, 24, 3
    seq = torch.randint(0, 21, (bb, n))
    msa = torch.randint(0, 21, (bb, s, n))
    extra = torch.randint(0, 21, (bb, 7, n))                 # never embedded by the reference (quirk Q11): only its presence matters
    mk, mmk = _masks(bb, s, n, 13)
    _, emk = _masks(bb, s, n, 14)                            # extra_msa_mask must have the MSA's shape (the stack embeds `msa`)
    fx = {"cfg": mcfg, "state": state, "inputs": {"seq": seq, "msa": msa, "mask": mk, "msa_mask": mmk, "extra_msa": extra, "extra_msa_mask": emk}}
    with torch.no_grad():
        fx["out_fp32"] = model(seq, msa, mask=mk, msa_mask=mmk, extra_msa=extra, extra_msa_mask=emk).distance
        model.double()
        fx["out_fp64"] = model(seq, msa, mask=mk, msa_mask=mmk, extra_msa=extra, extra_msa_mask=emk).distance
        model.float()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            fx["out_autocast_bf16"] = model(seq, msa, mask=mk, msa_mask=mmk, extra_msa=extra, extra_msa_mask=emk).distance
    torch.save(fx, os.path.join(OUT, "alphafold2_extra_msa.pt"))
    print("wrote alphafold2_extra_msa")


if __name__ == "__main__":
    main()
