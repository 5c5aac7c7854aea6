"""Reads sharded over GPUs: what has to cross ranks so that the result equals one process reading everything.

The search itself needs no exchange (reads are independent, the marker index is replicated).  Two reference
semantics are order-dependent over the whole input (microbe_census.py:328-367) and one result is a sum:

* ``-n``: the first `nreads` KEPT reads in file order are searched, and the too-short / low-quality / duplicate
  counters stop at the read that filled the quota -> all-gather of per-shard kept counts, then `shard_quota`;
* ``-d``: a read is a duplicate if an earlier kept read has the same sequence (either strand) -> `exchange_duplicates`:
  one all-to-all of (fingerprint, global index, passed-QC) routed by fingerprint, the owner marks every record behind
  the first passing read of its group, a second all-to-all returns the marks (`resolve_duplicates` is the same rule on
  one host with all passing fingerprints gathered: kept as the reference the exchange is tested against);
* the additive results (counters, per-family sums) -> one all-reduce of `SearchResult.counts_vector()`.

Shards are contiguous blocks of the read stream in rank order.  The pure functions below are tested on CPU against
the oracle (tests/test_host.py); `sharded_search` glues them to torch.distributed (NCCL on GPUs).
"""
import numpy as np


def shard_quota(kept_counts, nreads, rank):
    """Quota of kept reads rank `rank` has to search: -1 = the whole shard and the quota is not reached in it."""
    if nreads is None or nreads < 0:
        return -1
    before = int(sum(kept_counts[:rank]))
    remaining = nreads - before
    if remaining <= 0:
        return 0
    if remaining > kept_counts[rank]:
        return -1
    return int(remaining)


def _fp_keys(fp):
    """(n, 2) uint64 -> n structured keys that sort like (a, b)"""
    fp = np.ascontiguousarray(fp, dtype=np.uint64)
    return fp.view([("a", "<u8"), ("b", "<u8")]).reshape(-1)


def resolve_duplicates(codes, fps, first_index, passing_fps, passing_index):
    """New per-read verdicts of one shard under -d.

    codes / fps: verdicts (0 keep, 1 too short, 2 low quality) and fingerprints of the shard's reads, whose global
    indices are first_index .. first_index + n - 1; passing_fps / passing_index: fingerprints and global indices of
    the reads with verdict 0 of ALL shards.  Returns a copy of `codes` with 3 where the read is a duplicate."""
    codes = np.array(codes, dtype=np.uint8, copy=True)
    if len(codes) == 0 or len(passing_index) == 0:
        return codes
    pk = _fp_keys(passing_fps)
    order = np.lexsort((np.asarray(passing_index), pk["b"], pk["a"]))
    pk, pidx = pk[order], np.asarray(passing_index)[order]
    first = np.ones(len(pk), bool)
    first[1:] = pk[1:] != pk[:-1]
    uk, uidx = pk[first], pidx[first]                     # first passing read of every fingerprint
    lk = _fp_keys(fps)
    pos = np.searchsorted(uk, lk)
    pos[pos >= len(uk)] = len(uk) - 1
    found = uk[pos] == lk
    gidx = first_index + np.arange(len(codes), dtype=np.int64)
    dup = found & (uidx[pos] < gidx) & (codes != 1)
    codes[dup] = 3
    return codes


def _exchange_marks(a, b, gp, group):
    """Core of the -d exchange on tensors of one device: a, b = fingerprint halves (int64), gp = global index << 1 | passed-QC
    of this rank's long-enough reads.  Returns an int64 tensor, 1 where the read is a duplicate."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = a.device
    owner = torch.remainder(a & 0x7FFFFFFFFFFFFFFF, world)
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world)
    send = torch.stack([a[order], b[order], gp[order]], dim=1).contiguous()
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    in_split, out_split = [int(x) for x in counts.tolist()], [int(x) for x in recv_counts.tolist()]
    recv = torch.empty((sum(out_split), 3), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=out_split, input_split_sizes=in_split, group=group)
    # owner: order by (a, b, index) with three stable sorts, then the smallest passing index of every group
    ra, rb, rgp = recv[:, 0], recv[:, 1], recv[:, 2]
    perm = torch.argsort(rgp, stable=True)
    perm = perm[torch.argsort(rb[perm], stable=True)]
    perm = perm[torch.argsort(ra[perm], stable=True)]
    sa, sb, sgp = ra[perm], rb[perm], rgp[perm]
    n = sa.numel()
    dup_sorted = torch.zeros(n, dtype=torch.int64, device=dev)
    if n:
        new = torch.ones(n, dtype=torch.bool, device=dev)
        new[1:] = (sa[1:] != sa[:-1]) | (sb[1:] != sb[:-1])
        gid = torch.cumsum(new.to(torch.int64), 0) - 1
        big = torch.iinfo(torch.int64).max
        first_pass = torch.full((int(gid[-1]) + 1,), big, dtype=torch.int64, device=dev)
        first_pass.scatter_reduce_(0, gid, torch.where((sgp & 1) == 1, sgp >> 1, torch.full_like(sgp, big)), reduce="amin")
        dup_sorted = ((sgp >> 1) > first_pass[gid]).to(torch.int64)
    marks = torch.empty(n, dtype=torch.int64, device=dev)
    marks[perm] = dup_sorted
    back = torch.empty(a.numel(), dtype=torch.int64, device=dev)
    dist.all_to_all_single(back, marks, output_split_sizes=in_split, input_split_sizes=out_split, group=group)
    dup = torch.empty(a.numel(), dtype=torch.int64, device=dev)
    dup[order] = back
    return dup


def exchange_duplicates(codes, fps, first_index, group=None, device=None):
    """`-d` across ranks without any rank seeing all reads: every long-enough read sends (fingerprint, global
    index, passed-QC flag) to the rank that owns its fingerprint (a mod world) with ONE all-to-all, the owner sorts its
    records by (fingerprint, index) on its device and marks every record behind the first QC-passing one of its group,
    and the marks travel back with a second all-to-all.  Same verdicts as `resolve_duplicates`, 1 / world of its work per
    rank and no host-side sort.  Host arrays in, a copy of `codes` with 3 where the read is a duplicate out."""
    import torch
    dev = device or torch.device("cpu")
    codes = np.array(codes, dtype=np.uint8, copy=True)
    sel = np.flatnonzero(codes != 1)                                   # too short is decided before the duplicate test
    fp = np.ascontiguousarray(fps, dtype=np.uint64)[sel].view(np.int64).reshape(-1, 2)
    a = torch.from_numpy(fp[:, 0].copy()).to(dev)
    b = torch.from_numpy(fp[:, 1].copy()).to(dev)
    gp = torch.from_numpy(((first_index + sel.astype(np.int64)) << 1) | (codes[sel] == 0)).to(dev)    # index << 1 | passed
    dup = _exchange_marks(a, b, gp, group)
    codes[sel[dup.cpu().numpy() == 1]] = 3
    return codes


class _CudaView:
    """raw device pointer -> object torch.as_tensor() wraps without a copy"""
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 2}


def exchange_duplicates_device(engine, first_index, group=None, device=None):
    """The same with the reads where they are, in the engine's device memory: libmcx partitions the fingerprint records
    by owner (mcx_dedup_begin), sorts and marks what this rank owns (mcx_dedup_owner) and applies the marks that come
    back (mcx_dedup_finish); this function only moves the two buffers between ranks (NCCL all-to-all).  Returns the
    refreshed QC counters; `engine.exchange_ms` = device time of the whole exchange."""
    import torch
    import torch.distributed as dist
    dev_idx = (device.index if device is not None and device.index is not None else torch.cuda.current_device())
    if getattr(engine, "device", dev_idx) != dev_idx:
        raise RuntimeError("exchange_duplicates_device: the engine lives on GPU %d but the exchange tensors on GPU %d "
                           "(one process per GPU: create the engine on the rank's own device)" % (engine.device, dev_idx))
    dev = device or torch.device("cuda", dev_idx)
    world = dist.get_world_size(group)
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream.synchronize()
    e0.record(stream)
    d_send, counts = engine.dedup_begin(world, first_index)
    n_send = sum(counts)
    send = (torch.as_tensor(_CudaView(d_send, (n_send, 3), "<i8"), device=dev) if n_send
            else torch.zeros((0, 3), dtype=torch.int64, device=dev))
    cnt = torch.tensor(counts, dtype=torch.int64, device=dev)
    rcnt = torch.empty_like(cnt)
    dist.all_to_all_single(rcnt, cnt, group=group)
    out_split = [int(x) for x in rcnt.tolist()]
    m = sum(out_split)
    recv = torch.empty((m, 3), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send, output_split_sizes=out_split, input_split_sizes=counts, group=group)
    stream.synchronize()
    d_marks = engine.dedup_owner(recv.data_ptr() if m else 0, m)
    marks = (torch.as_tensor(_CudaView(d_marks, (m,), "|u1"), device=dev) if m else torch.zeros(0, dtype=torch.uint8, device=dev))
    back = torch.empty(n_send, dtype=torch.uint8, device=dev)
    dist.all_to_all_single(back, marks, output_split_sizes=counts, input_split_sizes=out_split, group=group)
    stream.synchronize()
    qc = engine.dedup_finish(back.data_ptr() if n_send else 0)
    e1.record(stream)
    stream.synchronize()
    engine.exchange_ms = e0.elapsed_time(e1)
    return qc


def sharded_round(engine, batch, first_index, remaining, filter_dups=False, group=None, device=None):
    """One round of a streamed sharded run: every rank pushes ITS batch of the round (global index of its first read =
    first_index; an empty batch when the input has run out for it), duplicates are settled across ranks (and against the
    reads kept in earlier rounds: the owners remember them), the -n quota still open (`remaining`, None = no limit) is
    shared out by an all-gather of the kept counts, and the rank searches its part.  Returns (local SearchResult -- NOT
    reduced --, reads sampled by all ranks in this round)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    engine.push(batch)
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    qc = None
    if filter_dups:
        if dev.type == "cuda":
            qc = exchange_duplicates_device(engine, first_index, group=group, device=dev)
        else:
            codes, fps = engine.qc_export(True)
            qc = engine.qc_import(exchange_duplicates(codes, fps, first_index, group=group, device=dev))
    if qc is None:
        qc = engine.qc()
    kept = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(kept, torch.tensor([qc["kept"]], dtype=torch.int64, device=dev), group=group)
    kept = [int(k) for k in kept]
    quota = shard_quota(kept, remaining, rank)
    res = engine.search(quota)
    total = sum(kept)
    return res, (total if remaining is None or remaining < 0 else min(total, remaining))


def allreduce_result(res, group=None, device=None):
    """sum the additive results over the ranks, in place"""
    import torch
    import torch.distributed as dist
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    v = torch.from_numpy(res.counts_vector()).to(dev)
    dist.all_reduce(v, group=group)
    res.load_counts_vector(v.cpu().numpy())
    return res


def sharded_search(engine, batch, first_index, nreads=None, filter_dups=False, group=None, device=None, push=None):
    """Search this rank's block of reads (already `set_params`-ed engine) and return the all-reduced SearchResult.

    `push` (optional) replaces `engine.push(batch)`, e.g. to push device-resident buffers.
    Without an initialised process group this is the single-GPU path (the engine's own -d is used then)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    if push is not None:
        push()
    else:
        engine.push(batch)
    if world == 1:
        return engine.search(-1 if nreads is None else nreads)
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    qc = None
    if filter_dups:
        if dev.type == "cuda":
            qc = exchange_duplicates_device(engine, first_index, group=group, device=dev)
        else:
            codes, fps = engine.qc_export(True)
            qc = engine.qc_import(exchange_duplicates(codes, fps, first_index, group=group, device=dev))
    if qc is None:
        qc = engine.qc()
    kept = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(kept, torch.tensor([qc["kept"]], dtype=torch.int64, device=dev), group=group)
    quota = shard_quota([int(k) for k in kept], nreads, rank)
    res = engine.search(quota)
    v = torch.from_numpy(res.counts_vector()).to(dev)
    dist.all_reduce(v, group=group)
    res.load_counts_vector(v.cpu().numpy())
    return res
