"""Reads sharded over GPUs: what has to cross ranks so that the result equals one process reading everything.

The search itself needs no exchange (reads are independent, the marker index is replicated).  Two reference
semantics are order-dependent over the whole input (microbe_census.py:328-367) and one result is a sum:

* ``-n``: the first `nreads` KEPT reads in file order are searched, and the too-short / low-quality / duplicate
  counters stop at the read that filled the quota -> all-gather of per-shard kept counts, then `shard_quota`;
* ``-d``: a read is a duplicate if an earlier kept read has the same sequence (either strand) -> all-gather of the
  fingerprints of the QC-passing reads, then `resolve_duplicates` (the first passing read of a fingerprint group is
  kept, every later long-enough read of the group is a duplicate);
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


def sharded_search(engine, batch, first_index, nreads=None, filter_dups=False, group=None, device=None, push=None):
    """Search this rank's block of reads (already `set_params`-ed engine) and return the all-reduced SearchResult.

    `push` (optional) replaces `engine.push(batch)`, e.g. to push device-resident buffers.
    Without an initialised process group this is the single-GPU path (the engine's own -d is used then)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    qc = push() if push is not None else engine.push(batch)
    if world == 1:
        return engine.search(-1 if nreads is None else nreads)
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    if filter_dups:
        codes, fps = engine.qc_export(True)
        ok = codes == 0
        mine = torch.from_numpy(np.concatenate([fps[ok].view(np.int64).reshape(-1, 2),
                                                (first_index + np.flatnonzero(ok)).astype(np.int64)[:, None]], axis=1)).to(dev)
        sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.shape[0]], dtype=torch.int64, device=dev), group=group)
        cap = int(max(int(s) for s in sizes))
        padded = torch.zeros((cap, 3), dtype=torch.int64, device=dev)
        padded[:mine.shape[0]] = mine
        parts = [torch.zeros((cap, 3), dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(parts, padded, group=group)
        allp = np.concatenate([p[:int(s)].cpu().numpy() for p, s in zip(parts, sizes)])
        codes = resolve_duplicates(codes, fps, first_index, allp[:, :2].view(np.uint64), allp[:, 2])
        qc = engine.qc_import(codes)
    kept = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(kept, torch.tensor([qc["kept"]], dtype=torch.int64, device=dev), group=group)
    quota = shard_quota([int(k) for k in kept], nreads, rank)
    res = engine.search(quota)
    v = torch.from_numpy(res.counts_vector()).to(dev)
    dist.all_reduce(v, group=group)
    res.load_counts_vector(v.cpu().numpy())
    return res
