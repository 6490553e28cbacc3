"""One-line digest of bench.py JSON lines: python tools/show_line.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        par = d.get("parity") or {}
        lay = d.get("solver_layout") or {}
        e2e = d.get("e2e") or {}
        cpu = d.get("cpu_baseline") or {}
        roof = d.get("roofline") or {}
        print(f"{d['config']['workload']} N={d['n_gpus']} {d.get('scaling')} P/gpu={d['config']['phases_per_gpu']}: {d['value']:.0f} cases/s {d['ms_per_step']:.3f} ms | e2e {e2e.get('value', 0):.0f} | "
              f"cpu {cpu.get('value')} | roofline {roof.get('kernel')} {roof.get('frac', 0):.3f} | parity {par.get('max_rel')} / lu {par.get('max_rel_vs_plain_lu')} idx {par.get('critical_index_match')} | "
              f"graph {lay.get('step_graph')} | launches {d.get('gpu_launches')} | " + str({k: round(v, 3) for k, v in d.get("stage_ms", {}).items()}))
    except Exception as e:
        print(f, "ERR", e)
        try:
            print(open(f.replace(".json", ".err")).read()[-1500:])
        except Exception:
            pass
