#!/usr/bin/env python
"""The host's ceiling for the end-to-end leg of bench.py: N processes (one per GPU), pinned buffers, the same
host->device and device->host byte mix as one ingest step (10 GB in, ~2.7 GB out per GPU by default), both directions at
once, NO kernels.  Prints one JSON line (rank 0): per-direction and aggregate GB/s, max over ranks.

    python tools/pcie_ceiling.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \\
        tools/pcie_ceiling.py

Modes measured: h2d alone, d2h alone, both at once; `ingest_GBps` = input bytes of all ranks / time of the concurrent mode -
the number bench.py's `e2e.value` cannot exceed."""
import argparse
import json
import os

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb-in", type=float, default=10.0)
    ap.add_argument("--gb-out", type=float, default=2.7)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--pieces", type=int, default=5, help="copies per direction per step (the streaming path copies in pieces)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n_in, n_out = int(args.gb_in * 1e9), int(args.gb_out * 1e9)
    h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    d_in = torch.empty(n_in, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(n_out, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    cur = torch.cuda.current_stream(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run(up: bool, down: bool, k: int) -> float:
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s_up.wait_stream(cur)
        s_dn.wait_stream(cur)
        for _ in range(k):
            for p in range(args.pieces):
                if up:
                    lo, hi = n_in * p // args.pieces, n_in * (p + 1) // args.pieces
                    with torch.cuda.stream(s_up):
                        d_in[lo:hi].copy_(h_in[lo:hi], non_blocking=True)
                if down:
                    lo, hi = n_out * p // args.pieces, n_out * (p + 1) // args.pieces
                    with torch.cuda.stream(s_dn):
                        h_out[lo:hi].copy_(d_out[lo:hi], non_blocking=True)
        cur.wait_stream(s_up)
        cur.wait_stream(s_dn)
        b.record()
        barrier()
        ms = a.elapsed_time(b) / k
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    run(True, True, 1)
    up_ms, dn_ms, both_ms = run(True, False, args.steps), run(False, True, args.steps), run(True, True, args.steps)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "gb_in_per_gpu": args.gb_in, "gb_out_per_gpu": args.gb_out, "steps": args.steps,
                          "h2d_alone_GBps_per_gpu": n_in / up_ms / 1e6, "d2h_alone_GBps_per_gpu": n_out / dn_ms / 1e6,
                          "both_ms_per_step": both_ms, "h2d_concurrent_GBps_per_gpu": n_in / both_ms / 1e6,
                          "aggregate_bytes_GBps": world * (n_in + n_out) / both_ms / 1e6,
                          "ingest_GBps": world * n_in / both_ms / 1e6, "cpus": os.cpu_count(),
                          "what": "pinned host <-> device copies only, both directions at once, all ranks at once, max over ranks; "
                                  "ingest_GBps bounds bench.py's e2e.value for this byte mix"}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
