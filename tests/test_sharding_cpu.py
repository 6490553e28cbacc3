"""Host-side logic of the multi-GPU path on CPU: shard bounds, first-max merge, and a world_size-2 gloo run in
which each rank evaluates its phase block (with the oracle standing in for the GPU stage) and the all-gathered
critical phase equals the unsharded scan."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT, golden_params, load_golden, oracle_model


def test_shard_bounds_cover_and_are_contiguous():
    from jacket_b200.distributed import shard_bounds
    for n in (1, 7, 36, 360, 4096, 65536):
        for ws in (1, 2, 3, 4, 8):
            b = [shard_bounds(n, ws, r) for r in range(ws)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(ws - 1))
            sizes = [h - l for l, h in b]
            assert max(sizes) - min(sizes) <= 1


def test_shard_times_match_reference_formula():
    from jacket_b200.distributed import shard_times
    from jacket_b200.morison import phase_times
    T, n = 9.4, 360
    full = phase_times(T, n)
    got = np.concatenate([shard_times(T, n, 4, r)[0] for r in range(4)])
    assert np.array_equal(got, full)


def test_merge_critical_first_max_rule():
    from jacket_b200.distributed import merge_critical
    assert merge_critical([1.0, 3.0, 3.0, 2.0], [0, 9, 5, 7]) == (3.0, 5)       # tie -> smaller global index
    assert merge_critical([float("nan"), 2.0], [0, 4]) == (2.0, 4)
    assert merge_critical([5.0], [-1]) == (None, -1)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from jacket_b200.distributed import allgather_critical, shard_times
    from oracle import jacket_oracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden("default3_airy")
    p = golden_params(g)
    m = oracle_model(g)
    n_total = 36
    t, lo = shard_times(p["T"], n_total, world, rank)
    out = orc.morison_phases(m, orc.AiryWave(p["H"], p["T"], p["d"], p["U_c"]), t, p["wave_dir"], p["current_dir"],
                             p["Cd"], p["Cm"], p["rho_water"])
    tab, crit = orc.phase_table(out, t, g["wave_omega"].item())
    val, idx = allgather_critical(tab[crit, 2], lo + crit)
    q.put((rank, val, idx))
    dist.destroy_process_group()


def test_gloo_world2_critical_phase_matches_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    res = [q.get(timeout=10) for _ in procs]
    g = load_golden("default3_airy")
    for _, val, idx in res:
        assert idx == int(g["scan36_critical"]) == 35
        assert abs(val - g["scan36_table"][35, 2]) < 1e-9 * val
