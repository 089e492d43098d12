#!/usr/bin/env python
"""One-screen summary of a bench.py JSON line (file argument or stdin)."""
import json
import sys

src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
for line in src:
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    print(f"n_gpus {d['n_gpus']}  value {d['value']:.4g} {d['unit']}  {d['ms_per_step']:.4f} ms/step  [{d['config'].get('workload')}]")
    r = d.get("roofline") or {}
    print(f"  roofline: frac {r.get('frac')}  canonical {r.get('step_canonical_frac')}  kernels {r.get('kernel_ms')}")
    e = d.get("e2e") or {}
    print(f"  e2e {e.get('value'):.4g}  ({e.get('ms_per_step')} ms/step)")
    sh = d.get("sharded_filter")
    if sh:
        print(f"  sharded: transport {sh['transport']}  parity: {sh['parity']}")
        print(f"           blocks {sh['blocks']} timed {sh['timed_seconds']:.3f} s  off-rank slots {sh['max_slots_received_from_peers_last_step']}")
    if d.get("large_filter"):
        lf = d["large_filter"]
        print(f"  large_filter {lf['value']:.4g}  {lf['ms_per_step']:.4f} ms  {lf['roofline']['kernel_ms']}  canonical {lf['roofline']['step_canonical_frac']:.3f}")
    for k, v in (d.get("extras") or {}).items():
        pl = v.get("pair_loop", {}).get("pairs_per_s")
        print(f"  {k:26s} {v['value']:.4g}  {v['ms_per_step']:.4f} ms  {v['roofline']['kernel_ms']}" + (f"  pairs/s {pl:.4g}" if pl else ""))
    c = d.get("cpu_baseline")
    if c:
        print(f"  cpu_baseline {c['value']:.4g} on {c['cores']} cores ({c['kind']})")
    print(f"  clocks {d.get('clocks')}  launches {d.get('gpu_launches')}")
