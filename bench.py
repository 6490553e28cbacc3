#!/usr/bin/env python
"""bench.py -- headline benchmark: phase load cases / second (Morison + FEM solve + post) on N B200.

One "step" = the whole analysis of one structure on every rank: deterministic assembly of K_ff, blocked Cholesky (factor
once), then the phase scan of this rank's phases (Morison -> nodal loads -> two DMMA triangular sweeps -> reactions / member
forces / utilisation -> per-phase table -> on-device first-max), finished by the cross-rank critical-phase reduction (NCCL
all-gather of one (value, index) pair per rank + table gather).

Workloads (BASELINE.json configs): c4_jacket10k (default; configs[3], the configuration the metric is quoted on: synthetic
16-leg x 104-bay jacket = 10,000 members / 19,968 free DOF, Airy wave -- the pinned model --, 4,096 phases), c3_jacket2k
(configs[2]: 1,976 members, 1,024 phases), c2_default3 (configs[1]: the default 3-leg jacket, 360 phases), c1_default3
(configs[0]: one phase), c5_ensemble (configs[4]: 4,096 sea states x 16 phases on the c3 jacket, one factor).
--scaling weak (default): rank r of N evaluates phases [r*P, (r+1)*P) of a P*N-phase scan (per-GPU work fixed);
--scaling strong: the workload's P phases are dealt over the N ranks (configs[3] "sharded across 8xB200").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--scaling weak|strong]      our arm
  python bench.py --impl reference ...                            the reference's CPU path (oracle port), host cores

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import os
import sys

if "--impl" in sys.argv and "reference" in sys.argv:
    # the reference arm uses every host core it can, whatever launcher started it (torch.distributed.run exports
    # OMP_NUM_THREADS=1): pin the BLAS / OpenMP pools to the core count BEFORE numpy is imported
    _n = str(os.cpu_count() or 1)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = _n

import argparse
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "phase_load_cases_per_sec"
UNIT = "cases/s"
WORKLOADS = {
    # name: (legs, bays, default phases per GPU)
    "c4_jacket10k": (16, 104, 4096),
    "c3_jacket2k": (8, 41, 1024),
    "c2_default3": (None, None, 360),
    "c1_default3": (None, None, 1),
    "c5_ensemble": (8, 41, 65536),     # configs[4]: 4096 sea states x 16 phases on the c3 jacket, one factor
}
# the critical phase of the named scans: (fixture under tests/golden, key) -- reference output (c2, c3) or the oracle table whose
# rows the reference pins (c4)
CRITICAL_FIXTURES = {("c4_jacket10k", 4096): ("c4_oracle_scan4096.npz", "critical"), ("c3_jacket2k", 1024): ("gen8x41_scan1024.npz", "scan1024_critical"),
                     ("c2_default3", 360): ("default3_airy.npz", "scan360_critical"), ("c2_default3", 36): ("default3_airy.npz", "scan36_critical")}
FP64_PEAK_TFLOPS = 37.1    # measured on this pool (profiles/r01_fp64_peaks.json): DMMA m8n8k4 issue peak; MEASURED_PEAKS.json has no FP64 entry
L2_BYTES = 126e6
PARITY_TOL = 1e-9


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)     # 20 x ~5 ms: long enough for a few clock samples and to average host jitter
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4_jacket10k", choices=list(WORKLOADS))
    ap.add_argument("--phases", type=int, default=0, help="phases (per GPU when weak, in total when strong; default: the workload's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--solver", default="banded", choices=["banded", "dense"])
    ap.add_argument("--ordering", default="rcm", choices=["rcm", "natural"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="phases in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gather-table", action="store_true", help="resident step at N > 1: also all-gather the per-phase table to every rank")
    ap.add_argument("--option", action="append", default=[], help="library option key=value (jk_set_option), repeatable")
    ap.add_argument("--per-step", action="store_true", help="debug: print every timed step's duration (ms) to stderr")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload, kernel):
    """dram read + write bytes of all launches of `kernel` in ONE step, from the newest ncu --set full summary under profiles/
    (written by tools/ncu_summary.py --traffic-json from a capture of exactly one step); None when no capture of this workload
    is committed -- never a constant in this file."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.isfile(path):
        return None, None
    with open(path) as f:
        data = json.load(f)
    entry = data.get(workload, {}).get(kernel)
    return (entry.get("bytes_per_step"), entry.get("source")) if entry else (None, None)


def shard(args, world, rank):
    """(phases of this rank, first phase, phases in total) under weak / strong scaling."""
    from jacket_b200.distributed import shard_bounds
    P = args.phases or WORKLOADS[args.workload][2]
    n_total = P * world if args.scaling == "weak" else P
    lo, hi = shard_bounds(n_total, world, rank)
    return hi - lo, lo, n_total


def workload_config(args, st, n_free, P_local, n_total, world):
    """The config block: key-identical for the GPU arm and the reference arm (the driver compares them)."""
    p_wave = "Airy (fallback) H=17.038 T=9.4 d=50 Uc=1.7, wave 38 deg, current 38 deg" if args.workload != "c5_ensemble" else \
        "Airy (fallback), sea states H~U[2,16] T~U[6,16] dir~U[0,360), d=50 Uc=1.7"
    return {"workload": args.workload, "members": int(st.n_members), "nodes": int(st.n_nodes), "free_dof": int(n_free),
            "phases_per_gpu": int(P_local), "phases_total": int(n_total), "wave": p_wave, "scaling": args.scaling,
            "parallelism": f"phase-shard x{world}",
            "step": "assemble + Cholesky factor + phase scan (Morison, nodal loads, 2 sweeps, post, reduce) + cross-rank critical-phase reduction"}


def ensemble_states(n_states, d=50.0, seed=20250101):
    """H~U[2,16] m, T~U[6,16] s, dir~U[0,360) deg; reject H/L > 0.142 or H/d > 0.78 (README.md:74-76)."""
    import jacket_b200 as jb
    rng = np.random.default_rng(seed)
    H, T, D = [], [], []
    while len(H) < n_states:
        h, t, w = rng.uniform(2, 16, 4096), rng.uniform(6, 16, 4096), rng.uniform(0, 360, 4096)
        k = jb.dispersion_wavenumbers(t, d)
        ok = (h * k / (2 * np.pi) <= 0.142) & (h / d <= 0.78)
        H.extend(h[ok]); T.extend(t[ok]); D.extend(w[ok])
    return np.array(H[:n_states]), np.array(T[:n_states]), np.array(D[:n_states])


def build_case(workload):
    import jacket_b200 as jb
    legs, bays, _ = WORKLOADS[workload]
    p = jb.AnalysisParams(wave_model="Airy")
    if legs is None:
        nodes, members, fixed, top = jb.create_default_3leg_jacket()
    else:
        nodes, members, fixed, top = jb.generate_jacket(legs, bays)
    st = jb.build_structure(nodes, members, fixed, top, p)
    wave = jb.RaschiiWave(p.H, p.T, p.d, p.U_c, "Airy", p.N_harm)
    return jb, st, wave, p


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def _cpu_phase_chunk(args):
    """Morison for a chunk of phases (worker process)."""
    (xyz, conn, sec_id, sections, fixed, top, wave_args, mor_kw, t) = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle import jacket_oracle as orc
    model = orc.Model(xyz, conn, sec_id, sections, fixed, top)
    wave = orc.AiryWave(*wave_args)
    out = orc.morison_phases(model, wave, t, **mor_kw)
    return out["nodal_forces"], out["total_drag"], out["total_inertia"], out["total_morison"]


class CpuReference:
    """Times the oracle (a NumPy port of the reference path) on the host: setup (elements, dense assembly, LU factor of
    K_ff -- scipy getrf, the factor half of numpy.linalg.solve) once, then per step a bounded sample of S phases: Morison
    (fanned over all cores, one process per phase shard), multi-RHS getrs, reactions, member forces.  Whole-workload
    throughput is extrapolated linearly in P (the reference's cost is exactly linear in the phase count, GUI.py:695-714):
    value = P / (t_setup + P/S * t_step).  For the sea-state ensemble a sample "phase" is one (sea state, phase) case with
    that state's own wave (`states`)."""

    def __init__(self, workload, P, sample, pool=None, states=None, n_phase=16):
        import multiprocessing as mp
        import scipy.linalg as sla
        from oracle import jacket_oracle as orc
        self.orc, self.sla = orc, sla
        jb, st, wave, p = build_case(workload)
        self.p, self.P, self.st = p, P, st
        xyz, conn, sec_id, _, sections = st.pack()
        secs = [(s.D_outer, s.t, s.rho_steel) for s in sections]
        fixed, top = st.indices(st.get_bottom_nodes()), st.indices(st.get_top_nodes())
        self.model = orc.Model(xyz, conn, sec_id, secs, fixed, top)
        self.wave = orc.AiryWave(p.H, p.T, p.d, p.U_c)
        self.cores = os.cpu_count() or 1
        self.S = min(P, sample if sample > 0 else max(8, min(self.cores, 32)))
        self.mor_kw = dict(wave_direction=p.wave_dir, current_direction=p.current_dir, Cd=p.Cd, Cm=p.Cm, rho_water=p.rho_water)
        self.geom = (xyz, conn.astype(np.int64), sec_id.astype(np.int64), secs, fixed, top)
        self.t_all = orc.phase_times(p.T, P)
        self.states, self.n_phase = states, n_phase
        # pool: worker processes forked by the caller before CUDA was initialised (run_ours), else forked here
        self.pool = pool if pool is not None else (mp.get_context("fork").Pool(min(self.cores, self.S)) if self.cores > 1 and self.S > 1 else None)
        t0 = time.perf_counter()
        self.fem = orc.FEM(self.model, p.E, p.nu)
        K = self.fem.K_global
        free = self.fem.free_dofs
        K_ff = K[np.ix_(free, free)]
        self.t_assemble = time.perf_counter() - t0
        t0 = time.perf_counter()
        self.lu = sla.lu_factor(K_ff, overwrite_a=True, check_finite=False)
        self.t_factor = time.perf_counter() - t0
        self.Krows = K[self.fem.fixed_dofs, :].copy()
        del K, K_ff
        self.fem._K = None
        self.n_step = 0
        self.last = None

    def sample_indices(self):
        return (np.arange(self.S) * max(1, self.P // self.S) + self.n_step) % self.P

    def _morison(self, idx):
        """nodal forces [S,Nn,3] and total_morison [S,3] of the sampled cases + their static loads."""
        orc, p = self.orc, self.p
        if self.states is None:
            t = self.t_all[idx]
            packs = [self.geom + ((p.H, p.T, p.d, p.U_c), self.mor_kw, c) for c in np.array_split(t, min(self.cores, len(idx))) if len(c)]
            inter, sw = self.fem.static_loads(p.wave_dir, p.F_axial, p.F_shear, p.M_moment, p.M_torsion, "calculated")
            inters = [inter] * len(idx)
        else:
            H, T, D = self.states
            packs, inters = [], []
            for c in idx:
                s, ph = divmod(int(c), self.n_phase)
                kw = dict(self.mor_kw, wave_direction=float(D[s]))
                packs.append(self.geom + ((float(H[s]), float(T[s]), p.d, p.U_c), kw, np.array([ph * T[s] / self.n_phase])))
                inter, sw = self.fem.static_loads(float(D[s]), p.F_axial, p.F_shear, p.M_moment, p.M_torsion, "calculated")
                inters.append(inter)
        res = self.pool.map(_cpu_phase_chunk, packs) if self.pool is not None else [_cpu_phase_chunk(a) for a in packs]
        return np.concatenate([r[0] for r in res]), np.concatenate([r[3] for r in res]), inters, sw

    def step(self, keep=False):
        """One bounded sample: S cases through Morison -> loads -> solve -> reactions -> member forces."""
        fem = self.fem
        idx = self.sample_indices()
        self.n_step += 1
        nodal, tm, inters, sw = self._morison(idx)
        F = np.concatenate([fem.load_matrix(nodal[i:i + 1], inters[i], sw) for i in range(len(idx))]) if self.states is not None \
            else fem.load_matrix(nodal, inters[0], sw)
        U = np.zeros_like(F)
        U[:, fem.free_dofs] = self.sla.lu_solve(self.lu, F[:, fem.free_dofs].T, check_finite=False).T
        R = U @ self.Krows.T - F[:, fem.fixed_dofs]
        mf = fem.member_forces(U, self.p.fy)
        if keep:
            self.last = dict(idx=idx, F=F, U=U, R=R.reshape(len(idx), -1, 6), mf=mf, total_morison=tm)
        return float(np.max(mf["utilization"])) + float(np.abs(R).max()) * 0 + float(np.abs(tm).max()) * 0

    def converged(self, steps=2):
        """The kept sample solved to convergence (iterative refinement with 80-bit residuals on the same LU factors): at
        20k DOF the plain LU result is only reproducible to ~1e-8 (tests/test_oracle_golden.py), parity is judged against this."""
        fem, last = self.fem, self.last
        U = last["U"].copy()
        for _ in range(steps):
            r = fem.residual_extended(U, last["F"])[:, fem.free_dofs]
            U[:, fem.free_dofs] += self.sla.lu_solve(self.lu, r.astype(np.float64).T, check_finite=False).T
        R = (U @ self.Krows.T - last["F"][:, fem.fixed_dofs]).reshape(U.shape[0], -1, 6)
        return U, R, fem.member_forces(U, self.p.fy)

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()

    def throughput(self, t_step):
        total = self.t_assemble + self.t_factor + (self.P / self.S) * t_step
        return self.P / total

    def describe(self, t_step):
        return (f"{self.S} of {self.P} phases per step (Morison over {min(self.cores, self.S)} processes, getrs+post threaded BLAS on {self.cores} cores) "
                f"= {t_step:.2f}s; setup once: elements+dense assembly {self.t_assemble:.1f}s, LU n={len(self.fem.free_dofs)} "
                f"{self.t_factor:.1f}s; extrapolated linearly to {self.P} phases")


def relmax(a, b):
    """max|a-b| / max|b| -- the parity metric of SURVEY 7 (hard part 3)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))


def fem_summary_columns(U, reactions, rows):
    """Columns 8..15 of the per-phase table from one phase's full results (run_analysis' log lines GUI.py:2027-2054)."""
    tr = np.linalg.norm(np.asarray(U).reshape(-1, 6)[:, :3], axis=1)
    util = np.asarray(rows)[:, 6]
    m = int(np.argmax(util))
    R = np.asarray(reactions)[:, :3].sum(axis=0)
    return np.array([tr.max(), float(np.argmax(tr)), util[m], float(m), np.asarray(rows)[m, 5], R[0], R[1], R[2]])


def parity_report(ref, fetch, table_rows, critical_index, critical_fixture):
    """GPU results of the sampled cases against the CPU port's (plain LU, as the reference solves) and against the converged
    solution of the same equations.  fetch(i) -> dict(U, reactions [n_fixed,6], rows [M,7]); table_rows(i) -> table row."""
    keys = ("Fx_max_kN", "Fy_max_kN", "Fz_max_kN", "My_max_kNm", "Mz_max_kNm", "von_mises_max_MPa", "utilization")
    last = ref.last
    Uc, Rc, mfc = ref.converged()
    worst = {"lu": 0.0, "converged": 0.0}
    for k, i in enumerate(last["idx"]):
        got = fetch(int(i))
        row = table_rows(int(i))
        for tag, U, R, mf in (("lu", last["U"], last["R"], last["mf"]), ("converged", Uc, Rc, mfc)):
            errs = [relmax(got["U"], U[k]), relmax(got["reactions"], R[k])]
            errs += [relmax(got["rows"][:, j], mf[key][k]) for j, key in enumerate(keys)]
            want = fem_summary_columns(U[k], R[k], np.stack([mf[key][k] for key in keys], axis=1))
            errs += [abs(row[8] - want[0]) / want[0], abs(row[10] - want[2]) / want[2], abs(row[12] - want[4]) / want[4], relmax(row[13:16], want[5:])]
            tm = last["total_morison"][k]
            errs += [abs(row[2] - np.linalg.norm(tm) / 1000) / (np.linalg.norm(tm) / 1000), relmax(row[5:8], tm / 1000)]
            worst[tag] = max(worst[tag], max(errs))
    return {"max_rel": worst["converged"], "max_rel_vs_plain_lu": worst["lu"], "tolerance": PARITY_TOL, "ok": bool(worst["converged"] < PARITY_TOL),
            "phases": [int(i) for i in last["idx"]], "fields": "U, reactions, 7 member fields, table columns 2,5-8,10,12-15",
            "critical_index": int(critical_index), "critical_index_match": None if critical_fixture is None else bool(critical_fixture == critical_index),
            "reference_of_max_rel": "converged solution of the reference's K_ff U = F (oracle LU + iterative refinement, 80-bit residuals); "
                                    "max_rel_vs_plain_lu is against one numpy.linalg.solve as the reference runs it (its own noise at 20k DOF is ~1e-8)"}


def critical_fixture(workload, n_total):
    entry = CRITICAL_FIXTURES.get((workload, n_total))
    if entry is None:
        return None
    path = os.path.join(ROOT, "tests", "golden", entry[0])
    return int(np.load(path)[entry[1]]) if os.path.isfile(path) else None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = max(1, args.gpus)
    P_local, _, n_total = shard(args, world, 0)
    states = None
    if args.workload == "c5_ensemble":
        states = ensemble_states(n_total // 16)
    ref = CpuReference(args.workload, n_total, args.cpu_sample, states=states)
    for _ in range(args.warmup):
        ref.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ref.step()
    t_step = (time.perf_counter() - t0) / max(1, args.steps)
    ref.close()
    val = ref.throughput(t_step)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * n_total / val, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, ref.st, len(ref.fem.free_dofs), P_local, n_total, world),
            "note": "reference CPU path = NumPy port of JacketAnalysisGUI_v2.py (oracle/jacket_oracle.py, pinned to the reference's golden "
                    "vectors); the reference itself is pure Python and cannot travel to the GPU box; the whole job's phases are timed on this host",
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ref.cores, "kind": "port", "sample": ref.describe(t_step),
                             "threads": {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS")}},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons under load.  Started BEFORE the warm-up, so that nvidia-smi's start-up (NVML
    attach) does not fall into the timed steps; every row carries nvidia-smi's own timestamp and stop() keeps the rows
    inside the timed window (all rows under load -- warm-up + timed -- if the window is shorter than the sampling period)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def wait_first(self, timeout=1.5):
        """Block until nvidia-smi has delivered its first row (its start-up is over) or the timeout expires."""
        t_end = time.time() + timeout
        while self.proc is not None and not self.rows and time.time() < t_end:
            time.sleep(0.01)

    @staticmethod
    def _stamp(text):
        import datetime
        try:
            return datetime.datetime.strptime(text, "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def stop(self, window=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            self.th.join(timeout=2)
        except Exception:
            pass
        rows, scope = list(self.rows), "warm-up + timed steps"
        if window is not None:
            inside = [r for r in rows if (self._stamp(r[0]) or -1.0) >= window[0] - 0.01 and (self._stamp(r[0]) or 1e30) <= window[1] + 0.01]
            if inside:
                rows, scope = inside, "timed steps"
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        hi = [c for c in sm if c >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(hi)) if hi else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm), "scope": scope}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class GpuRun:
    """Process-group set-up, step timing (CUDA events on the engine's stream, max over ranks) and the L2 policy shared by the
    phase-scan and the ensemble workloads."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- jacket_b200 has no CPU path")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device(f"cuda:{self.local_rank}")
        if self.world > 1:
            dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.flush_buf = None

    def options(self):
        return {kv.split("=")[0]: int(kv.split("=")[1]) for kv in self.args.option}

    def set_l2_policy(self, working_set_bytes):
        """Inputs much larger than L2 need no flush; otherwise a 256 MB buffer is overwritten between timed steps (outside the
        per-step event pairs)."""
        if working_set_bytes < 4 * L2_BYTES:
            self.flush_buf = self.torch.empty(256 << 20, dtype=self.torch.uint8, device=self.dev)
            return f"per-step working set ~{working_set_bytes / 1e6:.0f} MB < 4 x L2: a 256 MB buffer is overwritten between timed steps (steps timed one by one)"
        return f"per-step working set ~{working_set_bytes / 1e9:.1f} GB >> 126 MB L2: no flush needed"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, eng):
        torch = self.torch
        self.barrier()
        l0 = eng.launch_count()
        out = None
        with torch.cuda.stream(self.stream):
            if self.flush_buf is None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                marks = []
                e0.record(self.stream)
                for _ in range(steps):
                    out = fn()
                    if self.args.per_step:
                        marks.append(torch.cuda.Event(enable_timing=True)); marks[-1].record(self.stream)
                e1.record(self.stream)
                pairs = None
            else:
                pairs = []
                for _ in range(steps):
                    self.flush_buf.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(self.stream)
                    out = fn()
                    b.record(self.stream)
                    pairs.append((a, b))
        self.barrier()
        if pairs is None:
            total = e0.elapsed_time(e1)
            per = [b - a for a, b in zip([0.0] + [e0.elapsed_time(m) for m in marks][:-1], [e0.elapsed_time(m) for m in marks])]
        else:
            per = [a.elapsed_time(b) for a, b in pairs]
            total = float(sum(per))
        if self.args.per_step and per:
            print(f"[rank {self.rank}] step ms:", " ".join(f"{v:.2f}" for v in per), file=sys.stderr, flush=True)
        ms = torch.tensor([total], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item()), eng.launch_count() - l0, out

    def finish(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def kernel_table(stage, st, dims, sst, P, p, n_gauss=15, ensemble=False):
    """Per-stage roofline entries: algorithmic work of one launch (SURVEY 8d per-case figure x the P cases of the launch)
    over the stage's measured duration."""
    hbm_peak, peak_src = peaks()
    M, Nn = st.n_members, st.n_nodes
    n, nb, bw, NT, hb = dims["n_free_dof"], dims["tile"], dims["band_tiles"], dims["n_tiles"], dims["dof_half_bandwidth"]
    slab = max(8, sst.get("sweep_slab", 32) or 32)
    ldP = -(-P // slab) * slab
    nnzL, nnz_env = sst["nnz_L"], sst["nnz_L_min_envelope"]
    sweep_flops_alg = 2.0 * nnzL * P                    # one multiply-add per stored non-zero of L and right-hand side
    sweep_flops_exec = 0.5 * sst["sweep_flops_executed_per_case"] * ldP
    # Morison: FP64 instructions the kernel issues per Gauss point (SASS count): 25 for members below the lowest trough (drag-only
    # loop, closed-form inertia), 44 otherwise.  The restated algorithmic count is 2 flop per such instruction (the
    # shortest formulation known for GUI.py:633-659); SURVEY 8d's 170 flop/point (sincos = 40) counts a formulation the
    # kernel does not execute (per-point sincos hoisted into tables) and is reported for reference only.
    xyz_b, conn_b = st.pack()[0], st.pack()[1]
    zmax = np.maximum(xyz_b[conn_b[:, 0], 2], xyz_b[conn_b[:, 1], 2])
    amp = 8.0 if ensemble else abs(p.H / 2.0)
    f_sub = float(np.mean(zmax <= -amp * (1.0 + 1e-9)))
    mor_instr = (25.0 * f_sub + 44.0 * (1.0 - f_sub)) * n_gauss * M * P
    k = {}

    def add(name, ms, bound, achieved, peak, unit, **extra):
        if ms is None or ms <= 0:
            return
        k[name] = dict(ms=ms, bound=bound, unit=unit, peak=peak, achieved=achieved / (ms * 1e-3), frac=achieved / (ms * 1e-3) / peak, **extra)

    add("morison", stage.get("morison"), "fp64", 2.0 * mor_instr * 1e-12, FP64_PEAK_TFLOPS, "TFLOP/s",
        submerged_member_fraction=f_sub, alg_flop_per_gauss_point=2.0 * (25.0 * f_sub + 44.0 * (1.0 - f_sub)),
        survey_convention_tflops=170.0 * n_gauss * M * P * 1e-12 / (max(stage.get("morison") or 1e-9, 1e-9) * 1e-3),
        note="achieved = 2 flop x FP64 instructions issued per Gauss point (25 submerged / 44 general, SASS) -- an upper bound on useful work: dry points skip")
    for name in ("solve_fwd", "solve_bwd"):
        add(name, stage.get(name), "tensor", sweep_flops_alg * 1e-12, FP64_PEAK_TFLOPS, "TFLOP/s",
            executed=sweep_flops_exec * 1e-12 / (max(stage.get(name) or 1e-9, 1e-9) * 1e-3))
    add("post", stage.get("post"), "hbm", (8.0 * 6 * Nn + 56.0 * M + 48.0 * dims["n_fixed"]) * P * 1e-9, hbm_peak, "GB/s")
    add("rhs", stage.get("rhs"), "hbm", (48.0 * M + 8.0 * 6 * Nn) * P * 1e-9, hbm_peak, "GB/s")
    add("factor", stage.get("factor"), "tensor", n * float(hb) ** 2 * 1e-12, FP64_PEAK_TFLOPS, "TFLOP/s",
        note="latency chain of n pivots on two 8-CTA clusters, concurrent with the Morison stage")
    add("assemble", stage.get("assemble"), "hbm", (8.0 * NT * (bw + 1) * nb * nb + 1152.0 * M) * 1e-9, hbm_peak, "GB/s")
    if "solve_fwd" in k and "solve_bwd" in k:
        sweep_name = "k_sweep" if sst["tma_sweep"] else "k_slab_sweep"
        ms = 0.5 * (k["solve_fwd"]["ms"] + k["solve_bwd"]["ms"])
        k[sweep_name] = dict(ms=ms, launches_per_step=2, bound="tensor", unit="TFLOP/s", peak=FP64_PEAK_TFLOPS,
                             achieved=sweep_flops_alg * 1e-12 / (ms * 1e-3), executed=sweep_flops_exec * 1e-12 / (ms * 1e-3),
                             achieved_min_envelope=2.0 * nnz_env * P * 1e-12 / (ms * 1e-3), nnz_L=nnzL, nnz_L_min_envelope=nnz_env,
                             items_per_slab=sst["sweep_items"], rhs_per_cta=slab, ctas=ldP // slab)
        k[sweep_name]["frac"] = k[sweep_name]["achieved"] / FP64_PEAK_TFLOPS
        k[sweep_name]["frac_executed"] = k[sweep_name]["executed"] / FP64_PEAK_TFLOPS
        k[sweep_name]["frac_min_envelope"] = k[sweep_name]["achieved_min_envelope"] / FP64_PEAK_TFLOPS
    return k, peak_src


def pick_roofline(kernels, stage, step_ms, workload, peak_src):
    """Dominant kernel = largest exposed share of the step's critical path (the factorisation only counts for what the Morison +
    load stage does not hide)."""
    sweep_name = "k_sweep" if "k_sweep" in kernels else ("k_slab_sweep" if "k_slab_sweep" in kernels else None)
    front = (stage.get("morison") or 0.0) + max(stage.get("rhs") or 0.0, 0.0)
    exposed = {"morison": front, "post": stage.get("post") or 0.0, "factor": max(0.0, (stage.get("factor") or 0.0) - front)}
    if sweep_name:
        exposed[sweep_name] = 2 * kernels[sweep_name]["ms"]
    dom = max((kk for kk in exposed if kk in kernels), key=lambda kk: exposed[kk])
    d = kernels[dom]
    traffic, traffic_src = measured_traffic(workload, dom)
    if traffic is not None:
        traffic /= d.get("launches_per_step", 1)        # the capture covers one step; the roofline entry is per "launch" as counted here
    return {"kernel": dom, "bound": "tensor" if d["bound"] in ("tensor", "fp64") else "hbm", "achieved": d["achieved"], "peak": d["peak"],
            "unit": d["unit"], "frac": d["frac"], "traffic": traffic, "traffic_source": traffic_src,
            "executed": d.get("executed"), "frac_executed": d.get("frac_executed"), "frac_min_envelope": d.get("frac_min_envelope"),
            "launches_per_step": d.get("launches_per_step", 1), "ms_per_launch": d["ms"], "share_of_step": exposed[dom] / step_ms,
            "peak_source": ("FP64 pipe, DMMA m8n8k4 issue peak measured on this pool (profiles/r01_fp64_peaks.json); MEASURED_PEAKS.json has "
                            "no FP64 entry" if d["unit"] == "TFLOP/s" else peak_src),
            "note": "achieved = algorithmic work per launch (sweeps: 2*nnz(L)*P flops with the exact non-zero count of the factor in use; "
                    "frac_min_envelope uses the smallest-envelope ordering's count) / measured launch duration (CUDA events); executed = DMMA "
                    "flops issued; traffic = dram bytes per launch from the committed ncu --set full capture (null until captured)"}


def stage_timers(eng, step, n=3):
    """Per-stage CUDA-event timers of one step.  Event timers cannot be read out of a replayed CUDA graph, so they are taken
    from `n` extra steps with the graph switched off, after the timed region."""
    graph = eng.get_option("cuda_graph")
    if graph:
        eng.set_option("cuda_graph", 0)
    for _ in range(n):
        step()
    stage = eng.timings()
    if graph:
        eng.set_option("cuda_graph", 1)
    if stage.get("solve_fwd2", -1.0) > 0:       # split factor: the forward sweeps run as two groups of launches around the factor join
        stage["solve_fwd_first_parts"] = stage["solve_fwd"]
        stage["solve_fwd"] = stage["solve_fwd"] + stage["solve_fwd2"]
    return stage


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # CPU baseline (rank 0, N = 1): its worker processes are forked HERE, before CUDA is initialised, and sit idle during the
    # GPU measurements; the baseline itself is timed after them, so that nothing it leaves behind (threaded BLAS pools, 6 GB of
    # freed factor storage) shares the host with the launch loop of the timed steps.
    cpu_pool = None
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    if want_cpu and (os.cpu_count() or 1) > 1:
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        cpu_pool = mp.get_context("fork").Pool(min(cores, args.cpu_sample if args.cpu_sample > 0 else max(8, min(cores, 32))))

    run = GpuRun(args)
    torch = run.torch
    import jacket_b200 as jb
    from jacket_b200 import _lib as L
    from jacket_b200.distributed import sharded_phase_scan, shard_times

    P, lo, n_total = shard(args, world, rank)
    _, st, wave, p = build_case(args.workload)
    eng = jb.Engine(st, device=run.local_rank, stream=run.stream.cuda_stream, ordering=args.ordering, solver=args.solver, options=run.options())
    st._engine = eng
    E, G = p.E, p.E / (2 * (1 + p.nu))
    eng.set_supports(st.indices(st.get_bottom_nodes()))
    F_static = jb.static_load(st, p)
    eng.set_static_load(F_static)
    eng.set_wave(wave)
    eng.set_morison(np.deg2rad(90.0 - p.wave_dir), np.deg2rad(90.0 - p.current_dir), p.rho_water, p.Cd, p.Cm, 15)
    t_host, lo = shard_times(wave.T, n_total, world, rank)
    t_dev = torch.as_tensor(t_host, device=run.dev)
    dims = eng.dims()
    ldP = -(-P // 32) * 32
    l2_note = run.set_l2_policy(8.0 * ldP * (dims["n_pad"] + 7 * st.n_members + 6 * dims["n_fixed"]))

    def step_resident():
        """inputs already in HBM: assemble + factor + scan + cross-rank reduction"""
        # jk_step_dev: assemble + factorisation (side streams, concurrent with the Morison + load stage) + scan, one graph launch
        # The cross-rank exchange is the critical-phase reduction (one value / index pair per rank); every rank keeps its own shard of
        # the per-phase table in HBM.  --gather-table also all-gathers the table to every rank each step (round 1's resident step).
        return sharded_phase_scan(eng, wave, n_total, p.fy, rank, world, gather_table=(world > 1 and args.gather_table), t_dev=t_dev.data_ptr(),
                                  host_results=False, moduli=(E, G))

    def step_e2e():
        """host buffers in, host results out, through the public API: every rank gets the merged critical phase and its own
        shard of the table, rank 0 the whole table"""
        eng.set_static_load(F_static)
        return sharded_phase_scan(eng, wave, n_total, p.fy, rank, world, gather_table=True, t_host=t_host, moduli=(E, G))

    sampler = ClockSampler(run.local_rank)
    if rank == 0:
        sampler.start()                 # before the warm-up: its start-up cost stays out of the timed steps
        sampler.wait_first()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    w0 = time.time()
    ms_total, launches, out = run.timed(step_resident, args.steps, eng)
    clocks = sampler.stop(window=(w0, time.time())) if rank == 0 else None
    value = n_total * args.steps / (ms_total * 1e-3)
    crit_dev = int(out["critical_index"])

    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step_e2e()
        ms_e2e, _, out_e = run.timed(step_e2e, args.steps, eng)
        ms_e2e /= args.steps
        e2e = {"value": n_total / (ms_e2e * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(t_host.nbytes + F_static.nbytes),
               "d2h_bytes_per_step": int((P + (n_total if (rank == 0 and world > 1) else 0)) * L.TABLE_NCOL * 8 + 16 * world + 4),
               "ms_per_step": ms_e2e, "critical_index": int(out_e["critical_index"]),
               "returns": "every rank: merged critical (value, index) + its own table shard on the host; rank 0: the full table"}

    sst = eng.solver_stats()          # before the stage-timer steps switch the graph off
    stage = stage_timers(eng, step_resident)
    residual = eng.residual()

    cpu_baseline, parity = None, None
    if want_cpu:
        try:
            ref = CpuReference(args.workload, P, args.cpu_sample, pool=cpu_pool)
            ref.step()
            t0 = time.perf_counter()
            n_cpu = 2
            for k in range(n_cpu):
                ref.step(keep=(k == n_cpu - 1))
            t_step = (time.perf_counter() - t0) / n_cpu
            cpu_baseline = {"value": ref.throughput(t_step), "unit": UNIT, "cores": ref.cores, "kind": "port",
                            "sample": ref.describe(t_step)}
            # parity of the resident GPU results (last scan: all P phases) on the phases the CPU leg has just computed
            table_host, _ = eng.read_table(P)

            def fetch(i):
                got = eng.fetch_phase(i, U=True, reactions=True, rows=True)
                return got

            parity = parity_report(ref, fetch, lambda i: table_host[i], crit_dev, critical_fixture(args.workload, n_total))
            ref.close()
            del ref
        except Exception as e:  # a baseline failure must not hide the GPU number
            import traceback
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}",
                            "trace": traceback.format_exc()[-600:]}

    if rank == 0:
        kernels, peak_src = kernel_table(stage, st, dims, sst, P, p)
        step_ms = ms_total / args.steps
        roofline = pick_roofline(kernels, stage, step_ms, args.workload, peak_src)
        cfg = workload_config(args, st, dims["n_free_dof"], P, n_total, world)
        cfg["l2"] = l2_note
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg,
                "solver_layout": {"solver": args.solver, "ordering": args.ordering, "tile": dims["tile"], "band_tiles": dims["band_tiles"],
                                  "n_tiles": dims["n_tiles"], "dof_half_bandwidth": dims["dof_half_bandwidth"], "factor_chains": dims["n_chains"],
                                  "step_graph": sst["step_graph"], "graph_kernels": sst["graph_kernels"],
                                  "options": {eng.lib.jk_option_name(i).decode(): eng.get_option(eng.lib.jk_option_name(i).decode())
                                              for i in range(eng.lib.jk_option_count())}},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
                "stage_ms": {k: stage[k] for k in ("assemble", "factor", "wave_setup", "morison", "rhs", "solve_fwd", "solve_bwd", "post", "reduce", "scan_total") if k in stage},
                "stage_ms_note": "CUDA-event stage timers of 3 extra steps launched without the CUDA graph (events cannot be read out of a replayed graph)",
                "kernels": kernels, "critical_index": crit_dev, "rel_residual": residual}
        print(json.dumps(line), flush=True)
    run.finish()


def run_ensemble(args):
    """--workload c5_ensemble: assemble + factor once + ONE batch of n_states x 16 load cases per step (per GPU).  Sea states
    are dealt to the ranks (weak: 4,096 states per GPU; strong: 4,096 states in total).  The public call takes host arrays and
    returns the host table, so the step IS the end-to-end path."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu_pool = None
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    if want_cpu and (os.cpu_count() or 1) > 1:
        import multiprocessing as mp
        cpu_pool = mp.get_context("fork").Pool(min(os.cpu_count() or 1, args.cpu_sample if args.cpu_sample > 0 else 16))
    run = GpuRun(args)
    torch = run.torch
    import jacket_b200 as jb
    from jacket_b200 import _lib as L
    n_phase = 16
    cases_local, lo, cases_total = shard(args, world, rank)
    n_states, s_lo, n_states_total = cases_local // n_phase, lo // n_phase, cases_total // n_phase
    _, st, wave, p = build_case(args.workload)
    eng = jb.Engine(st, device=run.local_rank, stream=run.stream.cuda_stream, ordering=args.ordering, solver=args.solver, options=run.options())
    st._engine = eng
    H, T, D = ensemble_states(n_states_total)
    sl = slice(s_lo, s_lo + n_states)
    E, G = p.E, p.E / (2 * (1 + p.nu))
    eng.set_supports(st.indices(st.get_bottom_nodes()))
    dims = eng.dims()
    C = n_states * n_phase
    l2_note = run.set_l2_policy(8.0 * C * (dims["n_pad"] + 7 * st.n_members))

    def step():
        eng.assemble(E, G)
        eng.factor(overlap=True)
        return jb.ensemble_scan(st, H[sl], T[sl], D[sl], n_phase, d=p.d, U_c=p.U_c, current_direction=p.current_dir, Cd=p.Cd, Cm=p.Cm,
                                rho_water=p.rho_water, E=E, nu=p.nu, fy=p.fy, params=p, engine=eng)

    sampler = ClockSampler(run.local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    for _ in range(max(args.warmup, 3)):
        res = step()
    w0 = time.time()
    ms, launches, res = run.timed(step, args.steps, eng)
    clocks = sampler.stop(window=(w0, time.time())) if rank == 0 else None
    worst = torch.tensor([res.table[:, :, 10].max()], dtype=torch.float64, device=run.dev)
    if world > 1:
        run.dist.all_reduce(worst, op=run.dist.ReduceOp.MAX)        # ensemble-wide governing utilisation
    stage = eng.timings()
    cpu_baseline, parity = None, None
    if want_cpu:
        try:
            ref = CpuReference(args.workload, C, args.cpu_sample or 16, pool=cpu_pool, states=(H[sl], T[sl], D[sl]), n_phase=n_phase)
            ref.step()
            t0 = time.perf_counter()
            ref.step(keep=True)
            t_step = time.perf_counter() - t0
            cpu_baseline = {"value": ref.throughput(t_step), "unit": UNIT, "cores": ref.cores, "kind": "port",
                            "sample": ref.describe(t_step).replace("phases", "(sea state, phase) cases")}
            flat = res.table.reshape(-1, L.TABLE_NCOL)
            parity = parity_report(ref, lambda i: eng.fetch_phase(i, U=True, reactions=True, rows=True), lambda i: flat[i], -1, None)
            parity["critical_index"] = None
            ref.close()
        except Exception as e:
            import traceback
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e!r}",
                            "trace": traceback.format_exc()[-600:]}
    if rank == 0:
        val = cases_total * args.steps / (ms * 1e-3)
        step_ms = ms / args.steps
        sst = eng.solver_stats()
        kernels, peak_src = kernel_table(stage, st, dims, sst, C, p, ensemble=True)
        roofline = pick_roofline(kernels, stage, step_ms, args.workload, peak_src)
        cfg = workload_config(args, st, dims["n_free_dof"], C, cases_total, world)
        cfg.update(sea_states_per_gpu=n_states, phases_per_state=n_phase, l2=l2_note)
        line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg,
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": int(8 * (3 * n_states + 18 * st.n_nodes)),      # H, T, direction per sea state + static load + 2 direction loads
                        "d2h_bytes_per_step": int(C * L.TABLE_NCOL * 8 + 16 * n_states),       # table + critical phase and wave number per sea state "ms_per_step": step_ms,
                        "note": "the timed step takes host arrays (H, T, direction per sea state) and returns the host table: value == e2e by construction"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity,
                "stage_ms": {k: stage[k] for k in ("assemble", "factor", "morison", "rhs", "solve_fwd", "solve_bwd", "post", "reduce", "scan_total", "h2d", "d2h") if k in stage},
                "kernels": kernels, "governing_utilisation": float(worst.item())}
        print(json.dumps(line), flush=True)
    run.finish()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "c5_ensemble":
        run_ensemble(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
