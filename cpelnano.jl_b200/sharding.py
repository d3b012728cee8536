"""Sharding of estimation regions over ranks (one process per GPU) and the final result gather.

Regions are independent (the reference farms them out with Distributed.pmap, src/nanopore.jl:406), so each rank owns a
contiguous block of the region list — contiguous so that outputs stay in input order — balanced by the cost proxy
n_init·M_r·L_r.  No collective is on the data path; only the per-region result records are gathered.
"""
import numpy as np


def shard_bounds(costs, world_size):
    """Contiguous partition of len(costs) items into world_size blocks with near-equal total cost.
    Returns int64 array b[world_size+1]; rank k owns items b[k]:b[k+1]."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if world_size <= 1 or n == 0:
        return np.array([0, n] + [n] * (world_size - 1), dtype=np.int64)[:world_size + 1]
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    targets = cum[-1] * np.arange(1, world_size) / world_size
    cuts = np.searchsorted(cum, targets, side="left")
    # choose the nearer of the two neighbouring boundaries
    cuts = np.where((cuts > 0) & (np.abs(cum[np.maximum(cuts - 1, 0)] - targets) <= np.abs(cum[np.minimum(cuts, n)] - targets)), cuts - 1, cuts)
    b = np.concatenate([[0], np.clip(cuts, 0, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(b)


def region_costs(grp_off, read_off, n_init):
    return n_init * np.diff(np.asarray(grp_off)).astype(np.float64) * np.diff(np.asarray(read_off)).astype(np.float64)


def gather_arrays(local, group=None, dst=0, device=None, out_pinned=False):
    """The final result gather: per-rank result arrays (dict name → numpy array or CPU torch tensor, first axis =
    the rank's regions / rows) are concatenated in rank order on rank `dst`.  One tensor collective, no pickling: every rank packs
    its arrays into a single byte buffer (on `device` for NCCL — the buffer then travels GPU→GPU over NVLink — or on the CPU for
    gloo), the buffers are padded to the largest rank's size and gathered with torch.distributed.gather; `dst` copies each rank's
    bytes back to host memory and splits them.  With a single process it is the identity.  Returns dict of numpy arrays (None on
    the other ranks)."""
    import torch
    import torch.distributed as dist
    as_t = lambda v: v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))      # noqa: E731
    tens = {k: as_t(v).contiguous() for k, v in local.items()}
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return {k: t.numpy() for k, t in tens.items()}
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    keys = sorted(tens)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    nbytes = torch.tensor([tens[k].numel() * tens[k].element_size() for k in keys], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(nbytes) for _ in range(world)]
    dist.all_gather(sizes, nbytes, group=group)
    sizes = torch.stack(sizes).cpu().numpy()                                   # [world, n_keys] bytes
    pad = lambda n: (n + 15) & ~15                                            # noqa: E731  (16-byte aligned fields)
    totals = np.array([sum(pad(int(b)) for b in row) for row in sizes])
    cap = int(totals.max())
    buf = torch.empty(cap, dtype=torch.uint8, device=dev)
    o = 0
    for k in keys:
        b = tens[k].numel() * tens[k].element_size()
        if b:
            buf[o:o + b].copy_(tens[k].view(-1).view(torch.uint8), non_blocking=True)
        o += pad(b)
    parts = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(buf, parts, dst=dst, group=group)
    if rank != dst:
        return None
    host = [torch.empty(int(totals[r]), dtype=torch.uint8, pin_memory=out_pinned and dev.type == "cuda") for r in range(world)]
    for r in range(world):
        host[r].copy_(parts[r][:int(totals[r])], non_blocking=True)
    if dev.type == "cuda":
        torch.cuda.synchronize(dev)
    out = {}
    for j, k in enumerate(keys):
        pieces = []
        for r in range(world):
            o = sum(pad(int(b)) for b in sizes[r][:j])
            b = int(sizes[r][j])
            t = host[r][o:o + b].view(tens[k].dtype)
            pieces.append(t.view(-1, *tens[k].shape[1:]) if tens[k].dim() > 1 else t)
        out[k] = torch.cat(pieces).numpy()
    return out


def gather_records(local, group=None, dst=0, device=None):
    """Gather per-rank result arrays (dict of numpy arrays, row = region) on rank `dst`, concatenated in rank order
    (gather_arrays: one tensor collective — NCCL with `device`, gloo in the CPU tests)."""
    return gather_arrays(local, group=group, dst=dst, device=device)


class ResultGather:
    """gather_arrays with every buffer allocated once, for callers that gather a same-sized result record many times (bench.py's
    end-to-end steps): each rank's results live in ONE pinned host byte buffer (the library writes its outputs straight into views
    of it); run() sends it host → device, gathers over NCCL to `dst`, and `dst` copies every rank's bytes into its own pinned
    buffer.  bytes_moved() = (host→device, device→host) bytes of one run on this rank."""

    def __init__(self, nbytes, device, dst=0, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.dst, self.group = torch, dist, dst, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device)
        n = torch.tensor([int(nbytes)], dtype=torch.int64, device=self.device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n, group=group)
        self.sizes = [int(x.item()) for x in sizes]
        self.cap = max(self.sizes)
        self.send = torch.empty(self.cap, dtype=torch.uint8, device=self.device)
        self.parts = [torch.empty(self.cap, dtype=torch.uint8, device=self.device) for _ in range(self.world)] if self.rank == dst else None
        self.host = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in self.sizes] if self.rank == dst else None

    def run(self, local_pinned_u8):
        """local_pinned_u8: this rank's result bytes (pinned CPU uint8 tensor of the size given at construction)."""
        n = self.sizes[self.rank]
        self.send[:n].copy_(local_pinned_u8[:n], non_blocking=True)
        self.dist.gather(self.send, self.parts, dst=self.dst, group=self.group)
        if self.rank == self.dst:
            for r in range(self.world):
                self.host[r].copy_(self.parts[r][:self.sizes[r]], non_blocking=True)
        self.torch.cuda.synchronize(self.device)
        return self.host

    def bytes_moved(self):
        return self.sizes[self.rank], (sum(self.sizes) if self.rank == self.dst else 0)


def labelings_for(S, n1, LMAX, rng=None):
    """Group labelings of the unmatched test (src/grp_perm_tests.jl:236-247): all C(S, n1) combinations in lexicographic order
    when there are fewer than LMAX, otherwise the distinct members of LMAX uniform draws from 1:C(S, n1) — the drawn INDICES are
    turned into combinations by unranking (combinatorial number system), so nothing is enumerated when C(S, n1) is astronomically
    large (the reference collects the iterator lazily; 32 samples would be 6·10⁸ rows).  Returns (labelings [n, n1] int32 with
    1-based sample ids, exact)."""
    import math
    from itertools import combinations
    import numpy as np
    Lc = math.comb(S, n1)
    if Lc < LMAX:
        return np.array(list(combinations(range(1, S + 1), n1)), dtype=np.int32).reshape(-1, n1), True
    rng = rng or np.random.default_rng()
    # rand(1:L, LMAX): Python integers, because C(S, n1) may not fit in 64 bits
    def draw():
        if Lc < 2 ** 62:
            return int(rng.integers(0, Lc))
        return ((int(rng.integers(0, 2 ** 62)) << 62) | int(rng.integers(0, 2 ** 62))) % Lc      # 124 random bits, bias < 2^-60
    idx = sorted({draw() for _ in range(LMAX)})
    out = np.empty((len(idx), n1), dtype=np.int32)
    for row, rank in enumerate(idx):
        # lexicographic unranking: choose the smallest first element whose block of C(S - x, n1 - 1) combinations contains `rank`
        x, k = 1, n1
        for j in range(n1):
            while True:
                c = math.comb(S - x, k - 1)
                if rank < c:
                    break
                rank -= c
                x += 1
            out[row, j] = x
            x += 1
            k -= 1
    return out, False
