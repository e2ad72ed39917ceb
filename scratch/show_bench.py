import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"] if d.get("e2e") else None,
      "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
for k, v in d["ops"].items():
    print(f"{k:14s} avg_ms={v['avg_ms']:9.3f} share={v['share_of_step']:.3f} ach={v['achieved']:9.1f} {v['unit']:8s} frac={v['frac']:.3f}")
print(d["clocks"], "loss", d.get("loss_last_step"))
