"""Pretty-print the per-kernel table of a bench.py JSON line (stdin or file)."""
import json
import sys

src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
for line in src:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    if "kernels" not in d:
        print(line[:300])
        continue
    print(f"value {d['value']:.4g} {d['unit']}  e2e {d['e2e']['value']:.4g}  tokens/s {d.get('padded_tokens_per_sec', 0):.4g}  "
          f"ms/step {d['ms_per_step']:.2f}  launches {d['gpu_launches']}  clocks {d['clocks']}")
    for k, v in d["kernels"].items():
        print(f"  {k:16s} {v['ms_total']:9.2f} ms  n={v['launches']:5d}  share {v['share']:.3f}  "
              f"{v.get('achieved', 0):9.1f} {v.get('unit', ''):8s} frac {v.get('frac', 0):.3f}")
    if d.get("hyena_layer"):
        print("  hyena_layer", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["hyena_layer"].items() if k != "note"})
    if d.get("roofline"):
        print("  roofline", d["roofline"])
    if "cpu_baseline" in d:
        print("  cpu_baseline", d["cpu_baseline"])
    for name, ex in (d.get("extra_configs") or {}).items():
        if "error" in ex:
            print(f"  extra[{name}] ERROR {ex['error']}")
            continue
        r = ex.get("roofline") or {}
        print(f"  extra[{name}] {ex['value']:.4g} {ex['unit']}  ms/step {ex.get('ms_per_step', float('nan')):.2f}  roofline {r.get('kernel')} "
              f"{r.get('bound')} frac {r.get('frac', 0):.3f}" + (f"  conv share {ex['conv_share']:.3f}" if "conv_share" in ex else "")
              + (f"  e2e {ex['e2e']['value']:.4g}" if "e2e" in ex else ""))
        for k, v in (ex.get("kernels") or {}).items():
            print(f"      {k:16s} {v['ms_total']:9.2f} ms  share {v['share']:.3f}  {v.get('bound')} frac {v.get('frac') or 0:.3f}")
        if ex.get("hyena_layer"):
            print("      hyena_layer", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in ex["hyena_layer"].items() if k != "note"})
    if "gpu_eager_baseline" in d:
        print("  gpu_eager_baseline", d["gpu_eager_baseline"])
    if "sharding" in d:
        print("  sharding", d["sharding"])
