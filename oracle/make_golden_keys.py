"""TEST INFRASTRUCTURE ONLY — full state_dict contract of the UNMODIFIED reference Alphafold2 (alphafold2.py:470-628).

Run in the dev container (needs /root/reference):   python -m oracle.make_golden_keys
Writes tests/golden/alphafold2_state_keys.json: for two constructor configurations, every state_dict key with its shape
and its initialisation class ("zeros" / "ones" / "other"), so the drop-in's key set, shapes and zero/one-init parity are
pinned key by key (tests/test_cabi_loads.py).  `ipa_block.*` never appears: IPABlock is a third-party module the
reference does not vendor; oracle/ref_loader.py stubs it, so it contributes no parameters here.  A real reference
checkpoint carries those extra keys: the drop-in loads it with strict=False.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "alphafold2_state_keys.json")


def describe(model):
    out = {}
    for k, v in model.state_dict().items():
        kind = "other"
        if torch.is_floating_point(v) and v.numel() > 0:
            if bool((v == 0).all()):
                kind = "zeros"
            elif bool((v == 1).all()):
                kind = "ones"
        out[k] = {"shape": list(v.shape), "init": kind}
    return out


def main():
    ref = load_reference()
    cfgs = {"c1": dict(dim=128, depth=2, heads=4, dim_head=32),
            "angles": dict(dim=32, depth=1, heads=2, dim_head=16, predict_angles=True, extra_msa_evoformer_layers=1,
                           max_rel_dist=8, templates_dim=16)}
    res = {}
    for name, cfg in cfgs.items():
        torch.manual_seed(0)
        res[name] = {"cfg": cfg, "keys": describe(ref.Alphafold2(**cfg))}
        print(name, len(res[name]["keys"]), "keys")
    with open(OUT, "w") as f:
        json.dump(res, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()
