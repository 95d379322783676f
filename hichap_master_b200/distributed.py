"""Multi-GPU plumbing (one process per GPU, torch.distributed over NCCL for setup and data
exchange; the per-iteration allreduce inside the ICE loop is issued by the library itself on
the compute stream -- see hc_ice_csr_balance).

Genome-wide matrices are row-block sharded: every rank bins the pairs it
was handed into sorted keys, the ranks agree on row boundaries with ~equal numbers of stored
entries, exchange keys so that each rank owns the complete rows of its block, and reduce them
into a local symmetric CSR.  ICE then needs exactly one allreduce of the marginal vector per
iteration.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import kernels
from .device import PairColumns


def nccl_comm_from_process_group(device) -> int:
    """A NCCL communicator for the library's in-loop allreduce, bootstrapped through the
    existing torch.distributed process group (rank 0's unique id is broadcast)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        uid = torch.frombuffer(bytearray(kernels.nccl_unique_id()), dtype=torch.uint8).clone()
    backend = dist.get_backend()
    buf = uid.to(device) if backend == "nccl" else uid
    dist.broadcast(buf, src=0)
    return kernels.nccl_comm_init(bytes(buf.cpu().numpy().tobytes()), world, rank)


def row_cuts_from_counts(row_counts: np.ndarray, world: int):
    """Row boundaries (world+1) giving every rank ~the same number of stored entries."""
    from .shard import row_block_splits
    row_ptr = np.concatenate([[0], np.cumsum(row_counts, dtype=np.int64)])
    return row_block_splits(row_ptr, world)


def exchange_plan(sorted_rows_hist_local: np.ndarray, cuts):
    """Number of local keys destined to each rank, given the local per-row key counts."""
    c = np.concatenate([[0], np.cumsum(sorted_rows_hist_local, dtype=np.int64)])
    return [int(c[cuts[r + 1]] - c[cuts[r]]) for r in range(len(cuts) - 1)]


def build_row_block_csr(pairs: PairColumns, res: int, start, chrom_bins, nbins: int, cis_only=False):
    """Collective: returns (SymCsr of this rank's row block, cuts)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = pairs.device
    col_bits = kernels.key_col_bits(nbins)
    skeys, _, n_valid = kernels.pairs_to_sorted_keys(pairs, res, start, chrom_bins, nbins, cis_only)
    m = int(n_valid.item())
    skeys = skeys[:m]
    # per-row key counts: the keys are sorted, so the first key of every row is a binary search away (nbins searches
    # instead of a histogram over up to 2 G keys).  This bookkeeping, used to agree on the row boundaries, is the only
    # torch math on the path.
    edges = torch.arange(nbins + 1, dtype=torch.int64, device=dev) << col_bits
    pos = torch.searchsorted(skeys, edges)                       # first key with row >= r
    hist = pos[1:] - pos[:-1]
    # cut on DISTINCT keys per row (what a rank will store after reduce-by-key), not on pairs: the many duplicate
    # pairs next to the diagonal would otherwise skew the cuts (a key held by several ranks is counted once per rank)
    if m > 1:
        first = torch.ones(m, dtype=torch.int32, device=dev)
        first[1:] = (skeys[1:] != skeys[:-1]).to(torch.int32)
        cd = torch.zeros(m + 1, dtype=torch.int64, device=dev)
        torch.cumsum(first, 0, out=cd[1:])
        del first
        total = cd[pos[1:]] - cd[pos[:-1]]
        del cd
    else:
        total = hist.clone()
    dist.all_reduce(total)
    cuts = row_cuts_from_counts(total.cpu().numpy(), world)
    send = exchange_plan(hist.cpu().numpy(), cuts)
    send_t = torch.tensor(send, dtype=torch.int64, device=dev)
    recv_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv_t, send_t)
    recv = [int(x) for x in recv_t.cpu().numpy()]
    inbox = torch.empty(sum(recv), dtype=torch.int64, device=dev)
    dist.all_to_all_single(inbox, skeys.contiguous(), output_split_sizes=recv, input_split_sizes=send)
    # the received runs are sorted individually: one more radix sort merges them
    merged, free = kernels.sort_keys_u64(inbox, 2 * col_bits) if inbox.numel() > 1 else (inbox, None)
    nv = torch.tensor([merged.numel()], dtype=torch.int64, device=dev)
    row0, row1 = cuts[rank], cuts[rank + 1]
    row_ptr, col, cnt = kernels.keys_to_csr(merged, nv, col_bits, row1 - row0, scratch=free, row0=row0)
    return kernels.SymCsr(row_ptr, col, cnt, nbins, row0=row0), cuts
