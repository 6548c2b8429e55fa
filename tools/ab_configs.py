#!/usr/bin/env python
"""A/B driver for the `configs` block: python tools/ab_configs.py name ...  (name[:ENV=VAL,...]; name under build_variants/, or "tree")"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for spec in sys.argv[1:]:
    name, _, envs = spec.partition(":")
    env = dict(os.environ)
    if name != "tree":
        env["WBC_B200_LIB"] = os.path.join(ROOT, "build_variants", name, "libwbc_b200.so")
    for kv in filter(None, envs.split(",")):
        k, v = kv.split("=")
        env[k] = v
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5", "--no-cpu-baseline",
                        "--min-seconds", "0.5", "--e2e-seconds", "0.2", "--config-seconds", "0.4"], env=env, capture_output=True, text=True, timeout=400)
    try:
        d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
        print(f"{spec:40s} value {d['value']/1e6:7.2f} M  e2e {d['e2e']['value']/1e6:7.2f} | " +
              " | ".join(f"{c['config']} {c['steps_per_s']/1e6:.2f}" for c in d["configs"]), flush=True)
    except Exception as ex:
        print(spec, "FAILED", ex, p.stderr[-800:], flush=True)
