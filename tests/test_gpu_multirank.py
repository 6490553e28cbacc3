"""Two ranks on two GPUs over NCCL (skipped on a one-GPU box): the sharded scan's host-result and resident paths return
the unsharded scan's table and critical phase on every rank (first-maximum rule of GUI.py:717)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    try:
        _run(rank, world, port, q)
    except Exception:                     # the parent prints the child's traceback instead of a bare exit code
        import traceback
        q.put((rank, "ERR", traceback.format_exc()))
        raise


def _run(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import jacket_b200 as jb
    from jacket_b200.distributed import shard_times, sharded_phase_scan
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    ap = jb.AnalysisParams(wave_model="Airy")
    nodes, members, fixed, top = jb.generate_jacket(6, 9)
    st = jb.build_structure(nodes, members, fixed, top, ap)
    wave = jb.RaschiiWave(ap.H, ap.T, ap.d, ap.U_c, "Airy", ap.N_harm)
    n_total = 96 + 7                      # ragged: ranks get 52 and 51 phases
    ref = jb.phase_scan(st, wave, n_total, wave_direction=ap.wave_dir, current_direction=ap.current_dir, params=ap)      # engine device = LOCAL_RANK
    eng = ref.engine
    t_host, lo = shard_times(wave.T, n_total, world, rank)
    out = []
    for _ in range(2):                    # twice: the pinned staging buffer is reused
        a = sharded_phase_scan(eng, wave, n_total, ap.fy, rank, world, t_host=t_host)
        out.append((a["critical_index"], a["critical_value"], a["table"]))
    t_dev = torch.as_tensor(t_host, device=f"cuda:{rank}")
    b = sharded_phase_scan(eng, wave, n_total, ap.fy, rank, world, t_dev=t_dev.data_ptr(), host_results=False)
    torch.cuda.synchronize()
    q.put((rank, ref.critical_index, ref.table, out, int(b["critical_index"]), float(b["critical_value"]), b["table"].cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_nccl_world2_sharded_scan_matches_unsharded():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=300) for _ in procs]
    for r in res:
        assert r[1] != "ERR", r[2]
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    for rank, ci, table, out, bi, bv, btab in res:
        for (ai, av, atab) in out:
            assert ai == ci and av == table[ci, 2]
            assert np.array_equal(atab, table)
        assert bi == ci and bv == table[ci, 2]
        assert np.array_equal(btab[:, 2:], table[:, 2:]) and np.array_equal(btab[:, 0], table[:, 0])
