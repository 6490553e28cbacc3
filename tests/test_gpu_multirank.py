"""Two ranks: the sharded scan's host-result and resident paths return the unsharded scan's critical phase on every rank,
every rank's own shard of the table, and the full table on the gathering rank (first-maximum rule of GUI.py:717).
On two GPUs the exchange runs over NCCL; on a one-GPU box the same two-rank logic (ragged shards, merge, gather to rank 0)
runs with both engines on cuda:0 and gloo as the transport."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q, backend="nccl"):
    try:
        _run(rank, world, port, q, backend)
    except Exception:                     # the parent prints the child's traceback instead of a bare exit code
        import traceback
        q.put((rank, "ERR", traceback.format_exc()))
        raise


def _run(rank, world, port, q, backend):
    dev = rank if backend == "nccl" else 0
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(dev))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import jacket_b200 as jb
    from jacket_b200.distributed import shard_times, sharded_phase_scan
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{dev}"))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    ap = jb.AnalysisParams(wave_model="Airy")
    nodes, members, fixed, top = jb.generate_jacket(6, 9)
    st = jb.build_structure(nodes, members, fixed, top, ap)
    wave = jb.RaschiiWave(ap.H, ap.T, ap.d, ap.U_c, "Airy", ap.N_harm)
    n_total = 96 + 7                      # ragged: ranks get 52 and 51 phases
    ref = jb.phase_scan(st, wave, n_total, wave_direction=ap.wave_dir, current_direction=ap.current_dir, params=ap)      # engine device = LOCAL_RANK
    eng = ref.engine
    t_host, lo = shard_times(wave.T, n_total, world, rank)
    out = []
    for _ in range(3):                    # three times: the two pinned staging buffers are used in turn
        a = sharded_phase_scan(eng, wave, n_total, ap.fy, rank, world, t_host=t_host)
        out.append((a["critical_index"], a["critical_value"], a["table"], a["local_table"], a["offset"]))
    assert np.array_equal(out[0][3], out[2][3]) and out[0][3] is not out[1][3]      # a result survives the next call
    if backend == "nccl":
        t_dev = torch.as_tensor(t_host, device=f"cuda:{dev}")
        b = sharded_phase_scan(eng, wave, n_total, ap.fy, rank, world, t_dev=t_dev.data_ptr(), host_results=False)
        torch.cuda.synchronize()
        resident = (int(b["critical_index"]), float(b["critical_value"]), b["table"].cpu().numpy())
    else:
        resident = None
    out = [(ci, cv, None if tab is None else tab.copy(), loc.copy(), off) for ci, cv, tab, loc, off in out]
    q.put((rank, ref.critical_index, ref.table, out, resident))
    dist.barrier()
    dist.destroy_process_group()


def _two_ranks(backend):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if backend == "gloo" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, backend)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=300) for _ in procs]
    for r in res:
        assert r[1] != "ERR", r[2]
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    for rank, ci, table, out, resident in res:
        for (ai, av, atab, loc, off) in out:
            assert ai == ci and av == table[ci, 2]
            assert np.array_equal(loc, table[off:off + loc.shape[0]])              # every rank: its own shard
            if rank == 0:
                assert np.array_equal(atab, table)                                  # the gathering rank: the full table
            else:
                assert atab is None
        if resident is not None:
            bi, bv, btab = resident
            assert bi == ci and bv == table[ci, 2]
            assert np.array_equal(btab, table)


def test_nccl_world2_sharded_scan_matches_unsharded():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _two_ranks("nccl")


def test_two_ranks_on_one_gpu_over_gloo():
    _two_ranks("gloo")
