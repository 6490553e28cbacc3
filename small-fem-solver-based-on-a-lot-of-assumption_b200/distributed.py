"""Multi-GPU phase sharding.

Phases are independent units: rank r of g evaluates the contiguous block
[r*P_local, (r+1)*P_local) of a P_local*g-phase scan on its own GPU (geometry and the
factor of K are replicated -- identical on every rank, the assembly is deterministic).
The only exchange is the final critical-phase reduction: one (max total_kN, first index)
pair per rank through an NCCL all-gather (16 B per rank over NVLink), merged with the
reference's first-maximum rule (GUI.py:717), plus an optional gather of the per-phase
table.  torch.distributed is plumbing only (process group, NCCL / gloo transport).
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .morison import fill_phase_deg


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
    larger value wins, ties go to the smaller index."""
    import torch
    vals, idxs = pairs[:, 0], pairs[:, 1]
    best = vals.max()
    idx = torch.where(vals == best, idxs, torch.full_like(idxs, float("inf"))).min()
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


def sharded_phase_scan(engine, wave, n_total, fy, rank=0, world_size=1, group=None, gather_table=True,
                       t_dev=None, t_host=None, host_results=True):
    """One rank's part of an n_total-phase scan + the cross-rank critical-phase reduction.

    Returns dict(local_table (torch view on device), offset, critical_value, critical_index (global),
    table (global [n_total,16] on every rank if gather_table else None)).
    With t_dev (a device pointer holding this rank's times) nothing crosses PCIe before the reduction; with
    host_results=False nothing crosses it afterwards either (critical pair and gathered table stay device tensors and
    the call does not synchronise with the host).
    """
    import torch
    import torch.distributed as dist
    lo, hi = shard_bounds(n_total, world_size, rank)
    P = hi - lo
    host_table = host_crit = None
    if t_dev is not None:
        engine.phase_scan_dev(P, t_dev, fy)
    else:
        if t_host is None:
            t_host, _ = shard_times(wave.T, n_total, world_size, rank)
        if world_size == 1:
            host_table, host_crit = engine.phase_scan(t_host, fy)   # already copies table + critical index to the host
        else:
            engine.phase_scan_begin(t_host, fy)                     # the local table is only needed on the device
    table, val, idx = device_views(engine, P)
    if world_size == 1:
        # nothing to exchange: no all-gather, no merge kernels, and the table is copied to the host at most once
        if not host_results:
            return dict(local_table=table, offset=lo, critical_value=val.reshape(()), critical_index=(idx + lo).reshape(()),
                        table=table if gather_table else None)
        if host_table is None:
            host_table, host_crit = engine.read_table(P)
        fill_phase_deg(host_table, wave.omega)
        return dict(local_table=table, offset=lo, critical_value=float(host_table[host_crit, 2]), critical_index=int(host_crit) + lo,
                    table=host_table if gather_table else None)
    stream = torch.cuda.ExternalStream(engine.stream(), device=f"cuda:{engine.device}")
    with torch.cuda.stream(stream):
        if not host_results:
            cval, cidx = allgather_critical(val, idx + lo, group=group, to_host=False)
            full = _allgather_table(table, n_total, world_size, P, group) if gather_table else None
            return dict(local_table=table, offset=lo, critical_value=cval, critical_index=cidx, table=full)
        # host results: both all-gathers are queued behind the scan and their outputs go to ONE pinned staging buffer with
        # asynchronous copies.  The copies run on a torch-owned stream behind the engine's: the pinned block then only
        # remembers a stream that outlives the engine (freeing it after jk_destroy would otherwise touch a dead stream).
        pair = torch.stack([val.reshape(()), (idx + lo).reshape(()).to(torch.float64)])
        pairs = torch.empty(world_size * 2, dtype=torch.float64, device=pair.device)
        dist.all_gather_into_tensor(pairs, pair, group=group)
        n_tab = n_total * L.TABLE_NCOL if gather_table else 0
        pin = _pinned(world_size * 2 + n_tab)
        buf = _allgather_table(table, n_total, world_size, P, group) if gather_table else None
        cs = _copy_stream(pair.device)
        cs.wait_stream(stream)
        with torch.cuda.stream(cs):
            pin[:world_size * 2].copy_(pairs, non_blocking=True)
            if gather_table:
                pin[world_size * 2:].copy_(buf.reshape(-1), non_blocking=True)
        engine.read_critical(P)          # synchronises the engine's stream; also surfaces a failed factorisation
        cs.synchronize()                 # the copies (and with them the temporaries) are done before anything is freed
        host = pin.numpy()
        hp = host[:world_size * 2].reshape(world_size, 2)
        cval, cidx = merge_critical(hp[:, 0], hp[:, 1].astype(np.int64))
        full = None
        if gather_table:
            full = host[world_size * 2:].reshape(n_total, L.TABLE_NCOL).copy()     # the staging buffer is reused by the next scan
            fill_phase_deg(full, wave.omega)
    return dict(local_table=table, offset=lo, critical_value=cval, critical_index=cidx, table=full)


_PINNED = {}
_COPY_STREAMS = {}


def _copy_stream(device):
    """One torch-owned copy stream per device."""
    import torch
    key = str(device)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


def _pinned(n):
    """Reusable page-locked float64 staging buffer of at least n elements (one per process)."""
    import torch
    buf = _PINNED.get("buf")
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 1), dtype=torch.float64, pin_memory=True)
        _PINNED["buf"] = buf
    return buf[:n]


def _allgather_table(table, n_total, world_size, P, group=None):
    """All-gather the ranks' [P_r,16] tables into one [n_total,16] device tensor (NCCL, on the current stream)."""
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
