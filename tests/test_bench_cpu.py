"""bench.py's host-side pieces that need no GPU: the reference arm's CPU port on a small sample (checked against the
golden critical phase), its pre-forked worker pool, the clock sampler's timestamp filter and the reference-arm JSON line."""
import json
import os
import subprocess
import sys
import time

import numpy as np

from conftest import ROOT


def test_cpu_reference_with_preforked_pool_matches_serial():
    import multiprocessing as mp
    sys.path.insert(0, ROOT)
    import bench
    pool = mp.get_context("fork").Pool(2)
    a = bench.CpuReference("c2_default3", 36, 4, pool=pool)
    va = a.step()
    a.close()
    b = bench.CpuReference("c2_default3", 36, 4)                 # forks its own workers
    vb = b.step()
    b.close()
    assert va == vb and np.isfinite(va) and va > 0
    assert a.throughput(0.1) > 0 and "4 of 36 phases" in a.describe(0.1)


def test_clock_sampler_keeps_rows_inside_the_timed_window():
    sys.path.insert(0, ROOT)
    import bench
    s = bench.ClockSampler(0)
    now = time.time()

    def stamp(t):
        import datetime
        return datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]

    class _Done:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    class _Th:
        def join(self, timeout=None): pass

    s.proc, s.th = _Done(), _Th()
    s.rows = [[stamp(now - 5.0), "1200", "1965", "300", "Not Active", "Not Active", "Not Active", "Not Active"],
              [stamp(now - 0.5), "1965", "1965", "900", "Not Active", "Not Active", "Not Active", "Active"],
              [stamp(now - 0.4), "1950", "1965", "900", "Not Active", "Not Active", "Not Active", "Not Active"]]
    got = s.stop(window=(now - 1.0, now))
    assert got["samples"] == 2 and got["scope"] == "timed steps" and got["sm_mhz"] == 1957.5 and got["reasons"] == ["sw_power_cap"]
    s.proc, s.th = _Done(), _Th()
    got = s.stop(window=(now + 10.0, now + 11.0))                # nothing inside: falls back to every row under load
    assert got["samples"] == 3 and got["scope"] == "warm-up + timed steps"
    assert abs(bench.ClockSampler._stamp(stamp(now)) - now) < 2e-3


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OPENBLAS_NUM_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2_default3", "--steps", "1",
                          "--warmup", "0", "--cpu-sample", "4"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "cases/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["threads"]["OPENBLAS_NUM_THREADS"] == str(os.cpu_count())      # pinned by the arm itself, whatever the launcher exported
    # the config block has exactly the keys of the GPU arm's (the driver compares the two lines)
    assert set(line["config"]) == {"workload", "members", "nodes", "free_dof", "phases_per_gpu", "phases_total", "wave", "scaling", "parallelism", "step"}
    assert line["config"]["phases_total"] == 360 and line["scaling"] == "weak"


def test_reference_arm_covers_the_whole_job_at_n_gpus():
    """N > 1: the arm times the job's total phase count (weak: N x P, strong: P) -- same config block as the GPU arm at N."""
    env = dict(os.environ)
    for scaling, total in (("weak", 72), ("strong", 36)):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2_default3", "--steps", "1",
                              "--warmup", "0", "--cpu-sample", "4", "--gpus", "2", "--phases", "36", "--scaling", scaling],
                             capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
        assert out.returncode == 0, out.stderr[-2000:]
        line = json.loads(out.stdout.strip().splitlines()[-1])
        assert line["config"]["phases_total"] == total and line["config"]["phases_per_gpu"] == total // 2 and line["n_gpus"] == 2


def test_parity_report_on_cpu_results():
    """bench.parity_report with the CPU port standing in for the GPU fetches: zero error against itself, the converged
    solution within 1e-9 of the plain LU on a small model, critical-index bookkeeping."""
    sys.path.insert(0, ROOT)
    import bench
    ref = bench.CpuReference("c2_default3", 36, 4)
    ref.step(keep=True)
    last = ref.last
    keys = ("Fx_max_kN", "Fy_max_kN", "Fz_max_kN", "My_max_kNm", "Mz_max_kNm", "von_mises_max_MPa", "utilization")
    pos = {int(i): k for k, i in enumerate(last["idx"])}

    def fetch(i):
        k = pos[i]
        return dict(U=last["U"][k], reactions=last["R"][k], rows=np.stack([last["mf"][key][k] for key in keys], axis=1))

    def row(i):
        k = pos[i]
        r = np.zeros(16)
        tm = last["total_morison"][k]
        r[2] = np.linalg.norm(tm) / 1000; r[5:8] = tm / 1000
        r[8:16] = bench.fem_summary_columns(last["U"][k], last["R"][k], fetch(i)["rows"])
        return r

    rep = bench.parity_report(ref, fetch, row, 35, bench.critical_fixture("c2_default3", 36))
    ref.close()
    assert rep["max_rel_vs_plain_lu"] == 0.0 and rep["max_rel"] < 1e-9 and rep["ok"] and rep["critical_index_match"] is True
    assert bench.critical_fixture("c4_jacket10k", 4096) == 4044 and bench.critical_fixture("c3_jacket2k", 1024) == 1009
    assert bench.critical_fixture("c2_default3", 360) == 353 and bench.critical_fixture("c4_jacket10k", 100) is None
