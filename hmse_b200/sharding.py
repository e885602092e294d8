"""Multi-GPU protocol of the hot path, independent of how the local work is
computed so that it can be exercised with torch.distributed/gloo on CPU:

* stitch_cuts       byte-range shards -> the single-stream cut list.  Every rank resolves its
                    shard speculatively from offset 0, then the exit of rank r-1 (8 bytes) becomes
                    the entry of rank r and the shard is re-resolved incrementally; repeated until
                    no entry changes (chains converge within a few chunks, so normally one round).
* exchange_dedup    global exact dedup: records {digest, gid} go to owner = le32(digest) % world
                    with one all-to-all, the owner answers the smallest gid per digest with a
                    second all-to-all (north_star: "GPU hash table partitioned by digest prefix
                    with NCCL all-to-all over NVLink").
* exchange_lsh       global LSH bucketing: band b is owned by rank b % world.  Every rank sends each owner the
                    columns of its key matrix that the owner holds (one all-to-all); because shards are
                    contiguous and global ids ascend with the rank, the rows arrive in global id order, so the
                    owner's matrix [N_total][bands_owned] goes straight into the stable per-band sort
                    (hmse_lsh_buckets) with id = row (north_star: "partitioned by ... band prefix").
One process per GPU; `dev` is the device the collectives' tensors live on (cuda for NCCL, cpu
for gloo)."""
from __future__ import annotations

from typing import Callable, List, Tuple

import torch
import torch.distributed as dist


def _all_gather_i64(value: int, dev, group=None) -> List[int]:
    world = dist.get_world_size(group)
    mine = torch.tensor([value], dtype=torch.int64, device=dev)
    out = [torch.empty(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [int(t.item()) for t in out]


def stitch_cuts(resolve: Callable[[int], Tuple[object, int]], n_own: int, dev, group=None, max_rounds: int = 64):
    """resolve(entry) -> (cuts, exit_off) over the local shard (exit_off relative to the shard
    start; the next shard's entry is exit_off - n_own).  Returns (cuts, entry, rounds)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    entry = 0
    cuts, exit_off = resolve(entry)
    rounds = 0
    for rounds in range(1, max_rounds + 1):
        exits = _all_gather_i64(exit_off, dev, group)
        owns = _all_gather_i64(n_own, dev, group)
        new_entry = 0 if rank == 0 else max(0, exits[rank - 1] - owns[rank - 1])
        changed = int(new_entry != entry)
        if changed:
            entry = new_entry
            cuts, exit_off = resolve(entry)
        if sum(_all_gather_i64(changed, dev, group)) == 0:
            break
    else:
        raise RuntimeError("shard boundary resync did not converge in %d rounds" % max_rounds)
    return cuts, entry, rounds


def exchange_counts(counts: List[int], dev, group=None) -> List[int]:
    """counts[o] = records this rank sends to owner o  ->  recv[r] = records arriving from rank r."""
    world = dist.get_world_size(group)
    send = torch.tensor(counts, dtype=torch.int64, device=dev)
    recv = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send, group=group)
    return [int(x) for x in recv.tolist()]


def all_to_all_bytes(send: torch.Tensor, send_counts: List[int], recv_counts: List[int], width: int, group=None):
    """send: uint8 tensor of sum(send_counts)*width bytes grouped by destination rank."""
    recv = torch.empty(sum(recv_counts) * width, dtype=torch.uint8, device=send.device)
    dist.all_to_all_single(recv, send, [c * width for c in recv_counts], [c * width for c in send_counts], group=group)
    return recv


def exchange_dedup(records: torch.Tensor, counts: List[int], owner_resolve: Callable[[torch.Tensor, int], torch.Tensor],
                   group=None) -> torch.Tensor:
    """records: uint8[sum(counts)*40] grouped by owner.  owner_resolve(recv_records, m) -> int64[m]
    canonical gid per received record.  Returns int64[sum(counts)] replies in send order."""
    dev = records.device
    recv_counts = exchange_counts(counts, dev, group)
    recv = all_to_all_bytes(records, counts, recv_counts, 40, group)
    answers = owner_resolve(recv, sum(recv_counts))
    back = all_to_all_bytes(answers.view(torch.uint8), recv_counts, counts, 8, group)
    return back.view(torch.int64)


def owned_bands(bands: int, rank: int, world: int) -> List[int]:
    """Bands owned by `rank`: b % world == rank."""
    return list(range(rank, bands, world))


def exchange_lsh(keys: torch.Tensor, group=None) -> Tuple[torch.Tensor, List[int], int]:
    """keys: int64 [n_local, bands] (uint64 bit patterns) of this rank's chunks, in stream order.
    Returns (owned int64 [N_total, len(my_bands)], my_bands, id_base): row g of `owned` holds the keys of global
    chunk g for the bands this rank owns; id_base is the global id of this rank's first chunk."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = keys.device
    n, bands = int(keys.shape[0]), int(keys.shape[1])
    counts_all = _all_gather_i64(n, dev, group)
    id_base = sum(counts_all[:rank])
    n_total = sum(counts_all)
    mine = owned_bands(bands, rank, world)
    parts, send_counts = [], []
    for o in range(world):
        cols = keys[:, o::world].contiguous().view(-1)      # [n, bands_o] row-major
        parts.append(cols)
        send_counts.append(int(cols.numel()))
    send = torch.cat(parts) if parts else keys.new_empty(0)
    recv_counts = [c * len(mine) for c in counts_all]
    recv = torch.empty(sum(recv_counts), dtype=keys.dtype, device=dev)
    dist.all_to_all_single(recv, send, recv_counts, send_counts, group=group)
    return recv.view(n_total, len(mine)) if mine else recv.view(n_total, 0), mine, id_base


class TorchFabric:
    """The collectives global_delta_bases needs, spelled with torch.distributed (any backend: gloo on CPU in the tests,
    NCCL on GPUs).  The product's default is the C fabric (hmse_b200.ingest.CFabric: hmse_alltoallv / hmse_lsh_exchange
    inside the library); both move the same bytes in the same order."""

    def __init__(self, group=None, dev=None):
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev = dev or torch.device("cpu")

    def allgather4(self, a: int, b: int = 0, c: int = 0, d: int = 0):
        mine = torch.tensor([a, b, c, d], dtype=torch.int64, device=self.dev)
        out = [torch.empty(4, dtype=torch.int64, device=self.dev) for _ in range(self.world)]
        dist.all_gather(out, mine, group=self.group)
        return [t.tolist() for t in out]

    def lsh_exchange(self, keys: torch.Tensor):
        owned, _, id_base = exchange_lsh(keys, self.group)
        return owned, id_base

    def alltoallv(self, send: torch.Tensor, send_counts, elem_bytes: int):
        recv_counts = exchange_counts([int(c) for c in send_counts], self.dev, self.group)
        recv = all_to_all_bytes(send.contiguous().view(-1), [int(c) for c in send_counts], recv_counts, elem_bytes, self.group)
        return recv, recv_counts


def _as_u8(t: torch.Tensor) -> torch.Tensor:
    """Flat uint8 view of a tensor's bytes (empty tensors of odd strides included)."""
    if t.numel() == 0:
        return torch.empty(0, dtype=torch.uint8, device=t.device)
    return t.contiguous().view(-1).view(torch.uint8)


def global_delta_bases(fab, ops, keys: torch.Tensor, sel: torch.Tensor, n: int, id_base: int, bands: int, min_votes: int = 4):
    """Base selection of the L4 layer over the WHOLE stream (README.md:1556-1559: the LSH index a chunk probes is
    global), for a stream sharded by byte range.  `keys` int64 [m, bands]: band keys of this rank's first occurrences
    (chunks `sel` of its n local chunks, stream order); global first-occurrence ids ("u ids") follow rank order.
      1. band keys to the band owners (fab.lsh_exchange); the owners sort their bands over all M first occurrences and
         take the bucket heads (ops.heads);
      2. the heads travel back to the chunks' ranks (one all-to-all), columns re-interleaved;
      3. root flags locally (ops.votes pass 0), gathered from everybody; bases among ALL roots (ops.votes pass 1);
      4. bases on another rank: their u ids go to the owners, (length, global chunk id) and the bytes come back
         (ops.chunk_bytes serves the requests).
    fab: allgather4 / lsh_exchange / alltoallv (TorchFabric, or the C fabric).  ops: heads(owned) -> int32 [M, bo];
    votes(heads, ubase, min_votes, root_all=None); chunk_bytes(chunk indices) -> (packed uint8, lens int64).
    Returns dict(base_loc int64[n]: local chunk index of the base, n + e for fetched base e, -1 none;
                 ext_data uint8, ext_off int64[n_ext + 1], ext_gid int64[n_ext] global chunk ids of the fetched bases;
                 ubase, M).  Equal to oracle.delta_bases over the whole stream (tests: gloo on CPU, NCCL on GPUs)."""
    W = fab.world
    dev = keys.device
    m = int(sel.numel())
    owned, ubase = fab.lsh_exchange(keys)
    M, bo = int(owned.shape[0]), int(owned.shape[1])
    info = fab.allgather4(m, id_base)
    ms = [int(r[0]) for r in info]
    ubases = [sum(ms[:r]) for r in range(W)]
    if ubases[fab.rank] != ubase or sum(ms) != M:
        raise RuntimeError("LSH exchange and first-occurrence counts disagree")
    heads_o = ops.heads(owned) if (M and bo) else torch.empty(0, dtype=torch.int32, device=dev)
    recv, _ = fab.alltoallv(_as_u8(heads_o), [ms[r] * bo for r in range(W)], 4)
    heads = torch.empty(m, bands, dtype=torch.int32, device=dev)
    off = 0
    for o in range(W):
        cols = len(range(o, bands, W))
        if cols and m:
            heads[:, o::W] = recv[off:off + m * cols * 4].view(torch.int32).view(m, cols)
        off += m * cols * 4
    root_local = ops.votes(heads, ubase, min_votes, None)
    root_all, _ = fab.alltoallv(_as_u8(root_local.repeat(W)), [m] * W, 1)
    base_u = ops.votes(heads, ubase, min_votes, root_all)                  # global first-occurrence ids, -1 none
    has = base_u >= 0
    local = has & (base_u >= ubase) & (base_u < ubase + m)
    remote = has & ~local
    base_loc = torch.full((n,), -1, dtype=torch.int64, device=dev)
    if m:
        base_loc[sel[local]] = sel[(base_u[local] - ubase)]
    req = torch.unique(base_u[remote]) if m else base_u[:0]               # ascending = grouped by owner rank
    ends = torch.tensor([ubases[r] + ms[r] for r in range(W)], dtype=torch.int64, device=dev)
    own_r = torch.bucketize(req, ends, right=True)
    cnt = torch.bincount(own_r, minlength=W).tolist() if req.numel() else [0] * W
    got, gcnt = fab.alltoallv(_as_u8(req), cnt, 8)
    n_got = sum(gcnt)
    want_j = sel[(got[:n_got * 8].view(torch.int64) - ubase)] if n_got else sel[:0]    # my chunks the others asked for
    packed, lens_w = ops.chunk_bytes(want_j)
    meta = torch.stack([lens_w, want_j + id_base], 1).contiguous()
    back, _ = fab.alltoallv(_as_u8(meta), gcnt, 16)
    back = back[:req.numel() * 16].view(torch.int64).view(-1, 2) if req.numel() else torch.empty(0, 2, dtype=torch.int64, device=dev)
    ext_len, ext_gid = back[:, 0].contiguous(), back[:, 1].contiguous()
    per_req, pos = [], 0
    for c in gcnt:
        per_req.append(int(lens_w[pos:pos + c].sum()) if c else 0)
        pos += c
    ext_data, _ = fab.alltoallv(packed[:int(lens_w.sum())] if n_got else packed[:0], per_req, 1)
    ext_off = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(ext_len, 0)])
    if m and req.numel():
        base_loc[sel[remote]] = n + torch.searchsorted(req, base_u[remote])
    return dict(base_loc=base_loc, ext_data=ext_data, ext_off=ext_off, ext_gid=ext_gid, ubase=ubase, M=M, n_unique=m)


def bind_to_gpu_numa(device: int):
    """Pins the calling process to the CPUs local to `device` (sysfs local_cpulist of its PCI function), so that pinned
    host buffers allocated afterwards land on the GPU's NUMA node and host<->device copies do not cross sockets -
    with eight ranks streaming 12 GB per step each, remote pages make the inter-socket link the bottleneck.
    Returns a short description, or None when the topology is not visible (containers) or nothing usable is left."""
    import os
    import torch
    try:
        pr = torch.cuda.get_device_properties(device)
        bus = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        base = "/sys/bus/pci/devices/" + bus
        with open(base + "/local_cpulist") as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use or use == allowed:
            return None
        os.sched_setaffinity(0, use)
        node = None
        try:
            with open(base + "/numa_node") as f:
                node = int(f.read().strip())
        except Exception:  # noqa: BLE001
            pass
        return "gpu %d (%s): %d local cpus, numa node %s" % (device, bus, len(use), node)
    except Exception:  # noqa: BLE001
        return None
