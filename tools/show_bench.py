"""Compact view of one or more bench.py JSON lines (headline + ivf + c5 blocks)."""
import json, sys
def row(tag, e):
    r = e.get("roofline", {})
    ps = e.get("parity_sample", {})
    dv = e.get("device_vs_e2e", {})
    rec = e.get("recall_at_k_vs_exact_f32", {}).get("value")
    cpu = e.get("cpu_baseline", {}).get("value")
    print(f"{tag:34s} gpus {e.get('n_gpus')} value {e['value']:>12.0f} e2e {e['e2e']['value']:>12.0f} ms {e['ms_per_step']:7.3f} kern_ms {r.get('kernel_ms', 0):7.3f} "
          f"frac {r.get('frac', 0):.3f} parity {ps.get('ids_equal')}/{ps.get('dist_bits_equal')}{'/ties ' + str(ps['tie_classes_equal']) if 'tie_classes_equal' in ps else ''} dev=e2e {dv.get('ids_equal')}/{dv.get('dist_bits_equal')} "
          f"recall {rec if rec is None else round(rec, 4)} cpu {cpu if cpu is None else round(cpu)} fallback {e.get('fallback_queries_total')} launches {e.get('gpu_launches')} "
          f"clk {e.get('clocks', {}).get('sm_mhz')} {e.get('clocks', {}).get('reasons')}")
for f in sys.argv[1:]:
    for ln in open(f):
        ln = ln.strip()
        if not ln.startswith("{"):
            continue
        l = json.loads(ln)
        print(f"== {f}")
        row(l["metric"], l)
        for e in l.get("ivf", []):
            row("  " + e["metric"].replace("QPS ivf ", "ivf ").replace(" euclidean nlist=4096", ""), e)
        if "c5" in l:
            row("  c5 " + l["c5"]["metric"], l["c5"])
