"""torchrun --nproc-per-node P tests/parallel_check_multi_gpu.py : sharded trunk on P GPUs vs the single-GPU path and the oracle."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # repo root (this file lives in tests/)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import alphafold2_b200 as A  # noqa: E402
from alphafold2_b200.parallel import _GraphedTrunk, sharded_evoformer_forward  # noqa: E402
from oracle import evoformer_oracle as O  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ok = True
    for (d, H, dh, N, S, masked) in [(64, 2, 32, 48, 8, True), (64, 2, 32, 40, 8, False), (256, 8, 64, 128, 16, False), (256, 8, 64, 256, 128, True)]:
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
        # the CUDA-graph replay of the schedule must reproduce the eager schedule bit for bit, also on new input tensors
        gt = _GraphedTrunk(evo, None)
        xg, mg = gt(x.cuda(), m.cuda(), cu(mask), cu(msa_mask))
        xg2, mg2 = gt((x * 1.0).cuda(), (m * 1.0).cuda(), cu(mask), cu(msa_mask))
        torch.cuda.synchronize()
        gt.release()                                              # (teardown would do it too: parallel._hook_teardown)
        res = {"rank": rank, "world": world, "cfg": [d, H, dh, N, S, masked],
               "x_vs_single": (xs - x1).abs().max().item(), "m_vs_single": (ms - m1_).abs().max().item(),
               "x_scale": x1.abs().max().item(), "m_scale": m1_.abs().max().item(),
               "graph": "eager-fallback: " + gt.failed if gt.failed else "replayed",
               "graph_vs_eager": max((xg - xs).abs().max().item(), (mg - ms).abs().max().item(),
                                     (xg2 - xs).abs().max().item(), (mg2 - ms).abs().max().item())}
        if N <= 128 and rank == 0:
            rx, rm = O.evoformer({k: v.double() for k, v in st.items()}, "", x.double(), m.double(), H, 2, mask, msa_mask, chunk=16)
            res["x_vs_oracle_rel"] = ((xs.double().cpu() - rx).abs().max() / rx.pow(2).mean().sqrt()).item()
            res["m_vs_oracle_rel"] = ((ms.double().cpu() - rm).abs().max() / rm.pow(2).mean().sqrt()).item()
        good = res["x_vs_single"] <= 2e-3 * res["x_scale"] and res["m_vs_single"] <= 2e-3 * res["m_scale"] and res["graph_vs_eager"] == 0.0
        res["ok"] = bool(good)
        ok = ok and good
        print(json.dumps(res), flush=True)
    from alphafold2_b200 import parallel as par
    live = [ex for ex in par._PEER_ARENAS.values() if ex is not None]
    timeouts = any(ex.error() for ex in live)
    if rank == 0:
        print(json.dumps({"exchange": "peer-store kernel" if live else "nccl all_to_all", "arenas": len(par._PEER_ARENAS),
                          "barrier_timeouts": bool(timeouts)}), flush=True)
    ok = ok and not timeouts
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
