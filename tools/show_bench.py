#!/usr/bin/env python
"""One-screen summary of bench.py JSON lines: show_bench.py file.json ..."""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if d.get("impl") == "reference":
            print(f, "REFERENCE", round(d["value"], 1), d["unit"], d["cpu_baseline"]["sample"][:90])
            continue
        r = d["roofline"]
        print(f"{f}: N={d['n_gpus']} {d['ms_per_step'] * 1e3:8.2f} us/step  {d['value']:12.0f} frames/s  frac {r['frac']:.3f}  "
              f"e2e {d['e2e']['value']:9.0f} (pcie {d['e2e'].get('pcie_frac', 0):.2f})  parity_ok={d.get('parity_ok')}  [{d['launch']}]")
        c = d.get("c5")
        if c:
            print(f"    c5: {c['frames_total']} frames on {c['n_gpus']} GPU(s) in {c['total_ms']:.2f} ms = {c['frames_per_s']:.0f} frames/s, "
                  f"{c['ms_per_step'] * 1e3:.1f} us/micro-batch, frac/GPU {c['frac_of_hbm_peak_per_gpu']:.3f}")
        for k, v in (d.get("peers") or {}).items():
            if "error" in v:
                print("    peer", k, v["error"][:100]); continue
            print(f"    peer {k}: {v['ms_per_step'] * 1e3:.2f} us  frac {v['roofline']['frac']:.3f}  e2e {v['e2e']['value']:.0f}  cpu {v['cpu_baseline']['value']:.0f}  "
                  f"stock-torch-gpu {v['stock_torch_b200'].get('value', 0):.0f}  parity {v['parity'] and v['parity']['ok']}")
        for k, v in (d.get("other_workloads") or {}).items():
            if "error" in v:
                print("    ", k, v["error"][:100]); continue
            print(f"    {k}: {v['ms_per_step'] * 1e3:.2f} us  frac {v['whole_job_frac_of_hbm_peak']:.3f}  parity {v['parity'] and v['parity']['ok']}")
        if d.get("cpu_baseline"):
            print("    cpu_baseline", round(d["cpu_baseline"]["value"], 1), "stock_torch_b200", round(d.get("stock_torch_b200", {}).get("value", 0), 1))
