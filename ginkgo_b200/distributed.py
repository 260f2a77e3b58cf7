"""Multi-GPU plumbing: 1-D row partition of a CSR matrix over the ranks of a
torch.distributed process group (NCCL on B200s, gloo in the CPU tests).

This is the set-up half of the reference's experimental::distributed::Matrix
(core/distributed/matrix.cpp `read_distributed` + the RowGatherer / IndexMap construction):
from a rank's rows with GLOBAL column indices it derives
  * the sorted list of remote columns the rank references (its "ghosts"),
  * the local column numbering into the extended vector [owned | ghosts],
  * per-peer receive counts, and -- after one counts exchange and one index-list exchange --
    the per-peer send counts and the list of owned entries each peer needs.
Only torch tensor ops and torch.distributed collectives are used, so the very same code
runs on CPU tensors with gloo (tests/test_dist_cpu.py) and on CUDA tensors with NCCL.
The per-iteration exchanges themselves are done by the CUDA library's own NCCL
communicator on the compute stream (include/ginkgo_b200.h, b200_halo_exchange_*)."""
import torch
import torch.distributed as dist


def uniform_offsets(n, world):
    """Partition::build_from_global_size_uniform: contiguous, near-equal row ranges"""
    return [r * n // world for r in range(world + 1)]


def build_partition(col_idxs_global, offsets, rank, group=None):
    """-> dict(col_idxs_local, n_local, n_ghost, ghosts, recv_counts, send_counts, send_idx)"""
    world = len(offsets) - 1
    dev = col_idxs_global.device
    r0, r1 = offsets[rank], offsets[rank + 1]
    c = col_idxs_global.long()
    remote = (c < r0) | (c >= r1)
    ghosts = torch.unique(c[remote])  # sorted ascending => grouped by owning rank
    off_t = torch.tensor(offsets, dtype=torch.long, device=dev)
    bounds = torch.searchsorted(ghosts, off_t)
    recv_counts = (bounds[1:] - bounds[:-1]).contiguous()
    # every rank learns what every rank wants from every rank
    if world > 1:
        gathered = [torch.empty_like(recv_counts) for _ in range(world)]
        dist.all_gather(gathered, recv_counts, group=group)
        all_counts = torch.stack(gathered)  # [wanter, owner]
    else:
        all_counts = recv_counts.reshape(1, 1)
    send_counts = all_counts[:, rank].contiguous()
    # index lists: my ghost segments go to their owners, I receive what peers need from me
    rc, sc = recv_counts.tolist(), send_counts.tolist()
    need = [torch.empty(sc[p], dtype=torch.long, device=dev) for p in range(world)]
    if world > 1:
        ops = []
        b = bounds.tolist()
        for p in range(world):
            if p == rank:
                continue
            if rc[p] > 0:
                ops.append(dist.P2POp(dist.isend, ghosts[b[p]:b[p + 1]].contiguous(), p, group))
            if sc[p] > 0:
                ops.append(dist.P2POp(dist.irecv, need[p], p, group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
    send_idx = (torch.cat(need) - r0).to(torch.int32) if sum(sc) else \
        torch.empty(0, dtype=torch.int32, device=dev)
    n_local = r1 - r0
    local = torch.where(remote, n_local + torch.searchsorted(ghosts, c), c - r0)
    return dict(col_idxs_local=local.to(col_idxs_global.dtype), n_local=n_local,
                n_ghost=int(ghosts.numel()), ghosts=ghosts, recv_counts=recv_counts.cpu(),
                send_counts=send_counts.cpu(), send_idx=send_idx)


def halo_exchange_torch(x_ext, part, rank, group=None):
    """reference implementation of the halo exchange with torch.distributed only (used by the
    CPU tests and as the checker of the NCCL path): fills x_ext[n_local:]"""
    world = part["recv_counts"].numel()
    if world == 1:
        return
    n_local = part["n_local"]
    sc, rc = part["send_counts"].tolist(), part["recv_counts"].tolist()
    send = x_ext[part["send_idx"].long()]
    ops, so, ro = [], 0, 0
    recv_views = []
    for p in range(world):
        if sc[p] > 0:
            ops.append(dist.P2POp(dist.isend, send[so:so + sc[p]].contiguous(), p, group))
        if rc[p] > 0:
            buf = torch.empty(rc[p], dtype=x_ext.dtype, device=x_ext.device)
            recv_views.append((ro, buf))
            ops.append(dist.P2POp(dist.irecv, buf, p, group))
        so += sc[p]
        ro += rc[p]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for ro, buf in recv_views:
        x_ext[n_local + ro:n_local + ro + buf.numel()] = buf
