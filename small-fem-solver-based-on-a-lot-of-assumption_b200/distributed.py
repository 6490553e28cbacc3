"""Multi-GPU phase sharding.

Phases are independent units: rank r of g evaluates the contiguous block [lo_r, hi_r) of an n_total-phase scan on its
own GPU (geometry and the factor of K are replicated -- identical on every rank, the assembly is deterministic).  There
is no data-path collective.  The only exchange is the final critical-phase reduction: one (max total_kN, first global
index) pair per rank through an NCCL all-gather (16 B per rank over NVLink), merged with the reference's first-maximum
rule (GUI.py:717), plus -- on request -- a gather of the per-phase table to ONE rank.  torch.distributed is plumbing only
(process group, NCCL / gloo transport).

Host results: every rank gets the merged critical pair and its OWN shard of the table; the full table is assembled on
`table_rank` only (default 0; None: nowhere).  Nothing is broadcast back and no rank touches another rank's rows on the
host, so the per-step host work does not grow with the number of ranks.  Tables returned as host arrays are views of two
alternating page-locked buffers: a result stays valid until the second-next call on the same engine.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L


def shard_bounds(n_total, world_size, rank):
    """Contiguous block of rank ``rank``: sizes differ by at most one, earlier ranks take the remainder."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_times(T, n_total, world_size, rank):
    """t_i = i*T/n_total for the rank's block, evaluated exactly like GUI.py:696."""
    lo, hi = shard_bounds(n_total, world_size, rank)
    return np.array([i * T / n_total for i in range(lo, hi)], dtype=np.float64), lo


def merge_critical(values, indices):
    """First maximum over ranks: larger value wins, ties go to the smaller global index; NaN never wins."""
    best_v, best_i = None, -1
    for v, i in zip(values, indices):
        if i < 0 or v != v:
            continue
        if best_v is None or v > best_v or (v == best_v and i < best_i):
            best_v, best_i = float(v), int(i)
    return best_v, best_i


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned device memory."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def device_views(engine, P):
    """torch views (no copy) of the table [P,16], the critical value [1] and index [1] of the last scan."""
    import torch
    lib = engine.lib
    ptrs = (lib.jk_table_dev(engine.h), lib.jk_critical_value_dev(engine.h), lib.jk_critical_index_dev(engine.h), P)
    cached = engine.__dict__.get("_dev_views")
    if cached is not None and cached[0] == ptrs:          # the library's buffers only move when they grow
        return cached[1]
    dev = f"cuda:{engine.device}"
    table = torch.as_tensor(_DevArray(ptrs[0], (P, L.TABLE_NCOL), "<f8"), device=dev)
    val = torch.as_tensor(_DevArray(ptrs[1], (1,), "<f8"), device=dev)
    idx = torch.as_tensor(_DevArray(ptrs[2], (1,), "<i8"), device=dev)
    engine.__dict__["_dev_views"] = (ptrs, (table, val, idx))
    return table, val, idx


def merge_critical_device(pairs):
    """First-maximum merge of an all-gathered [world, 2] (value, global index) tensor ON THE DEVICE (no host sync):
    larger value wins, ties go to the smaller index; NaN pairs (a rank whose factorisation failed) never win."""
    import torch
    vals, idxs = pairs[:, 0], pairs[:, 1]
    ok = (vals == vals) & (idxs >= 0)
    best = torch.where(ok, vals, torch.full_like(vals, float("-inf"))).max()
    idx = torch.where(ok & (vals == best), idxs, torch.full_like(idxs, float("inf"))).min()
    return best, idx


def allgather_critical(local_value, local_global_index, group=None, device=None, to_host=True):
    """All-gather one (value, global index) pair per rank and merge.  Works on NCCL (device tensors) and gloo (CPU).
    to_host=False keeps the merged pair on the device (two 0-d tensors), so the step stays asynchronous."""
    import torch
    import torch.distributed as dist
    if isinstance(local_value, torch.Tensor):
        pair = torch.stack([local_value.reshape(()).to(torch.float64), local_global_index.reshape(()).to(torch.float64)])
    else:
        pair = torch.tensor([float(local_value), float(local_global_index)], dtype=torch.float64, device=device or "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        ws = dist.get_world_size(group)
        out = torch.empty(ws * 2, dtype=torch.float64, device=pair.device)   # flat: gloo and NCCL both accept it
        dist.all_gather_into_tensor(out, pair, group=group)
        out = out.reshape(ws, 2)
    else:
        out = pair.reshape(1, 2)
    if not to_host:
        return merge_critical_device(out)
    out = out.cpu().numpy()
    return merge_critical(out[:, 0], out[:, 1].astype(np.int64))


def gather_table_to(table, n_total, world_size, rank, group=None, dst=0):
    """Gather the ranks' [P_r,16] tables to rank ``dst`` as one [n_total,16] tensor (None elsewhere).  NCCL or gloo; runs on
    the current stream.  Shards are padded to the largest block so that one flat gather serves ragged splits too."""
    import torch
    import torch.distributed as dist
    sizes = [shard_bounds(n_total, world_size, r) for r in range(world_size)]
    pmax = max(h - l for l, h in sizes)
    mine = table.contiguous()
    if mine.shape[0] < pmax:
        mine = torch.cat([mine, mine.new_zeros((pmax - mine.shape[0], L.TABLE_NCOL))])
    parts = [torch.empty_like(mine) for _ in range(world_size)] if rank == dst else None
    dist.gather(mine, parts, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([p[:h - l] for p, (l, h) in zip(parts, sizes)]) if any(h - l < pmax for l, h in sizes) else torch.cat(parts)


def _allgather_table(table, n_total, world_size, P, group=None):
    """All-gather the ranks' [P_r,16] tables into one [n_total,16] device tensor on EVERY rank (resident path)."""
    import torch
    import torch.distributed as dist
    sizes = [shard_bounds(n_total, world_size, r) for r in range(world_size)]
    if all(h - l == P for l, h in sizes):
        buf = torch.empty(n_total * L.TABLE_NCOL, dtype=torch.float64, device=table.device)
        dist.all_gather_into_tensor(buf, table.contiguous().reshape(-1), group=group)
        return buf.reshape(n_total, L.TABLE_NCOL)
    parts = [torch.empty((h - l, L.TABLE_NCOL), dtype=torch.float64, device=table.device) for l, h in sizes]
    dist.all_gather(parts, table.contiguous(), group=group)
    return torch.cat(parts)


def sharded_phase_scan(engine, wave, n_total, fy, rank=0, world_size=1, group=None, gather_table=True,
                       t_dev=None, t_host=None, host_results=True, table_rank=0, moduli=None):
    """One rank's part of an n_total-phase scan + the cross-rank critical-phase reduction.

    Returns dict(local_table, offset, critical_value, critical_index (global), table).
      host_results=True   critical pair as Python numbers; local_table = this rank's shard [P_r,16] (host);
                          table = the full [n_total,16] host table on rank `table_rank` if gather_table, None elsewhere.
      host_results=False  everything stays on the device and the call does not synchronise with the host: critical pair as
                          0-d tensors, local_table = view of the library's table, table = all-gathered device tensor on
                          every rank if gather_table.
    With t_dev (device pointer to this rank's times) nothing crosses PCIe before the reduction either.
    moduli=(E, G): the call is a whole step -- K is re-assembled and re-factored first (jk_step / jk_step_dev: one library
    call, replayed as a CUDA graph); without it the engine's current factor is used.
    """
    import torch
    import torch.distributed as dist
    lo, hi = shard_bounds(n_total, world_size, rank)
    P = hi - lo
    if t_dev is not None:
        if moduli is None:
            engine.phase_scan_dev(P, t_dev, fy)
        else:
            engine.step_dev(moduli[0], moduli[1], P, t_dev, fy)
    else:
        if t_host is None:
            t_host, _ = shard_times(wave.T, n_total, world_size, rank)
        if world_size == 1 and host_results:                    # one pinned copy of table + critical index, one synchronisation
            table, crit = engine.phase_scan(t_host, fy) if moduli is None else engine.step(moduli[0], moduli[1], t_host, fy)
            return dict(local_table=table, offset=lo, critical_value=float(table[crit, 2]), critical_index=int(crit) + lo,
                        table=table if gather_table else None)
        if moduli is None:
            engine.phase_scan_begin(t_host, fy)                 # queue only: the table is needed on the device first
        else:
            engine.step(moduli[0], moduli[1], t_host, fy, read=False)
    table, val, idx = device_views(engine, P)
    # everything below is ordered behind the scan on the engine's own (non-blocking) stream
    stream = torch.cuda.ExternalStream(engine.stream(), device=f"cuda:{engine.device}")
    caller = torch.cuda.current_stream(torch.device(f"cuda:{engine.device}"))
    with torch.cuda.stream(stream):
        if not host_results:
            if world_size == 1:
                out = dict(local_table=table, offset=lo, critical_value=val.reshape(()).clone(), critical_index=(idx + lo).reshape(()),
                           table=table if gather_table else None)
            else:
                cval, cidx = allgather_critical(val, idx + lo, group=group, to_host=False)
                full = _allgather_table(table, n_total, world_size, P, group) if gather_table else None
                out = dict(local_table=table, offset=lo, critical_value=cval, critical_index=cidx, table=full)
            # the returned tensors (views of the library's buffers among them) are consumed on the caller's stream: order it
            # behind the engine's stream (an event wait on the device, no host synchronisation)
            if caller.cuda_stream != stream.cuda_stream:
                caller.wait_stream(stream)
            return out
        if world_size == 1:                                     # device times in, host results out
            host_table, host_crit = engine.read_table(P)
            return dict(local_table=host_table, offset=lo, critical_value=float(host_table[host_crit, 2]), critical_index=int(host_crit) + lo,
                        table=host_table if gather_table else None)
        # host results on several ranks: pair all-gather + table gather to one rank are queued behind the scan; their outputs
        # and this rank's shard go to ONE pinned staging buffer with asynchronous copies on a torch-owned stream (the pinned
        # block then only remembers a stream that outlives the engine), and the host synchronises once.
        want_full = gather_table and table_rank is not None
        n_loc = P * L.TABLE_NCOL
        n_full = n_total * L.TABLE_NCOL if (want_full and rank == table_rank) else 0
        pin = _pinned(engine, world_size * 2 + n_loc + n_full)
        pair = torch.stack([val.reshape(()), (idx + lo).reshape(()).to(torch.float64)])
        cs = _copy_stream(pair.device)
        if dist.get_backend(group) == "gloo":
            # CPU transport (tests on a box with fewer GPUs than ranks): this rank's pair and shard come to the host first,
            # the collectives run on host tensors
            cs.wait_stream(stream)
            with torch.cuda.stream(cs):
                pin[:2].copy_(pair, non_blocking=True)
                pin[world_size * 2:world_size * 2 + n_loc].copy_(table.reshape(-1), non_blocking=True)
            engine.read_critical(P)
            cs.synchronize()
            mine = pin[:2].clone()
            dist.all_gather_into_tensor(pin[:world_size * 2], mine, group=group)
            if want_full:
                full_cpu = gather_table_to(pin[world_size * 2:world_size * 2 + n_loc].reshape(P, L.TABLE_NCOL), n_total, world_size, rank, group, table_rank)
                if full_cpu is not None:
                    pin[world_size * 2 + n_loc:].copy_(full_cpu.reshape(-1))
        else:
            pairs = torch.empty(world_size * 2, dtype=torch.float64, device=pair.device)
            dist.all_gather_into_tensor(pairs, pair, group=group)
            full_dev = gather_table_to(table, n_total, world_size, rank, group, table_rank) if want_full else None
            cs.wait_stream(stream)
            with torch.cuda.stream(cs):
                pin[:world_size * 2].copy_(pairs, non_blocking=True)
                pin[world_size * 2:world_size * 2 + n_loc].copy_(table.reshape(-1), non_blocking=True)
                if full_dev is not None:
                    pin[world_size * 2 + n_loc:].copy_(full_dev.reshape(-1), non_blocking=True)
            engine.read_critical(P)          # synchronises the engine's stream; also surfaces a failed factorisation
            cs.synchronize()                 # the copies (and with them the temporaries) are done before anything is freed
        host = pin.numpy()
        hp = host[:world_size * 2].reshape(world_size, 2)
        cval, cidx = merge_critical(hp[:, 0], hp[:, 1].astype(np.int64))
        local = host[world_size * 2:world_size * 2 + n_loc].reshape(P, L.TABLE_NCOL)
        full = host[world_size * 2 + n_loc:].reshape(n_total, L.TABLE_NCOL) if n_full else None
    return dict(local_table=local, offset=lo, critical_value=cval, critical_index=cidx, table=full)


_COPY_STREAMS = {}


def _copy_stream(device):
    """One torch-owned copy stream per device."""
    import torch
    key = str(device)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


def _pinned(engine, n):
    """Page-locked float64 staging: two buffers per engine used in turn, so the arrays handed out by one call stay intact
    during the next call (no copy out of the staging area)."""
    import torch
    state = engine.__dict__.setdefault("_pinned_pair", {"bufs": [None, None], "turn": 0})
    k = state["turn"]
    state["turn"] = 1 - k
    buf = state["bufs"][k]
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 1), dtype=torch.float64, pin_memory=True)
        state["bufs"][k] = buf
    return buf[:n]
