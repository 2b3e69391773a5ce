#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line."""
import json
import sys

d = json.load(open(sys.argv[1]))
print("%s  N=%d  value %.3f G/s  %.1f us/step   e2e %.3f G/s (%.1f us/step)" % (
    d["config"]["workload"][:3], d["n_gpus"], d["value"] / 1e9, d["ms_per_step"] * 1e3, d["e2e"]["value"] / 1e9, d["e2e"]["ms_per_step"] * 1e3))
c16 = d["e2e"].get("compact16")
if c16:
    print("   e2e over 16-B records: %.3f G/s (%.1f us/step)" % (c16["value"] / 1e9, c16["ms_per_step"] * 1e3))
r = d.get("roofline")
if r:
    print("   replay stage %.1f us  frac %.4f  whole step frac %.4f  traffic %s  step traffic/alg %s" % (
        r["avg_launch_ms"] * 1e3, r["frac"], r.get("whole_step_frac", 0), r.get("traffic"), r.get("step_traffic_over_algorithmic")))
    if r.get("traffic_per_kernel"):
        print("   DRAM bytes per launch:", {k: round(v / 1e6, 2) for k, v in r["traffic_per_kernel"].items()}, "MB")
c = d.get("cpu_baseline")
if c:
    print("   cpu_baseline:", {k: c[k] for k in c if k != "sample"})
for k, x in (d.get("extra") or {}).items():
    if "error" in x:
        print("   extra", k, "ERROR", x["error"])
        continue
    rr = x.get("roofline") or {}
    print("   extra %s: %.3f G/s  %.1f us/step  parity %s/%s  k_main frac %s  whole-step frac %s  imbalance %s" % (
        k, x["value"] / 1e9, x["ms_per_step"] * 1e3, x["parity"]["gpu_verdict_mismatches"], len(x["parity"]["table_mismatch_ranks"]),
        round(rr.get("frac", 0), 4), round(rr.get("whole_step_frac", 0), 4), (x.get("imbalance") or {}).get("owner_load_max_over_mean")))
print("   clocks", d.get("clocks"), " launches", d.get("gpu_launches"), " hot_rows", d.get("hot_rows"))
