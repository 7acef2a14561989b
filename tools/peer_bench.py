"""torchrun --nproc-per-node P tools/peer_bench.py : time of one row<->column exchange of the pair tensor, peer-store kernel
(every AF2_PEER_VARIANT) against NCCL all_to_all_single + its pack/unpack copies, at the C2 and C4 shapes."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alphafold2_b200 import parallel as par  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    for (N, S, d, iters) in [(256, 128, 256, 200), (1024, 512, 256, 20)]:
        ex = par.peer_exchange_for(None, dev, N, S, d)
        shard_mb = N // world * N * d * 4 / 1e6
        res = {"N": N, "world": world, "shard_MB": round(shard_mb, 1), "remote_MB": round(shard_mb * (world - 1) / world, 1)}
        if ex is not None:
            xr = ex.buffer(par.PeerExchange.PAIR, "row")
            xr.normal_()

            def both():
                xc = ex.rows_to_cols(xr, par.PeerExchange.PAIR)
                ex.cols_to_rows(xc, par.PeerExchange.PAIR)
            for v in ("2", "0", "6"):
                os.environ["AF2_PEER_VARIANT"] = v
                ms = timed(both, iters) / 2
                res[f"peer_v{v}_us"] = round(ms * 1e3, 1)
                res[f"peer_v{v}_remote_GBs"] = round(res["remote_MB"] / ms, 1)
            os.environ["AF2_PEER_VARIANT"] = "2"
            ref = xr.clone()
            both()
            res["roundtrip_exact"] = bool(torch.equal(ref, xr))
            res["barrier_timeouts"] = ex.error()
        t = torch.randn(N // world, N, d, device=dev)

        def both_nccl():
            par.cols_to_rows(par.rows_to_cols(t, None), None)
        ms = timed(both_nccl, iters) / 2
        res["nccl_us"] = round(ms * 1e3, 1)
        res["nccl_remote_GBs"] = round(res["remote_MB"] / ms, 1)
        if rank == 0:
            print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
