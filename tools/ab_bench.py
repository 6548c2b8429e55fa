#!/usr/bin/env python
"""A/B driver for the GPU box: runs bench.py (no configs, no CPU baseline) once per variant and prints one line each.
usage: python tools/ab_bench.py name[:ENV=VAL[,ENV=VAL]] ...   (name = directory under build_variants/, or "tree" for the in-tree library)"""
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
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "5", "--no-cpu-baseline", "--no-configs",
                        "--min-seconds", "0.8", "--e2e-seconds", "0.4"], env=env, capture_output=True, text=True, timeout=300)
    try:
        d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
        e = d["e2e"]
        print(f"{spec:40s} value {d['value']/1e6:7.2f} M  e2e {e['value']/1e6:7.2f} M ({e['host_path']})  cand "
              + " ".join(f"{k}={v/1e6:.1f}" for k, v in e['host_path_candidates_steps_per_s'].items())
              + f"  open_res {e['open_loop_resident_state']['value']/1e6:6.2f}  all_host {e['all_inputs_from_host']['value']/1e6:6.2f}"
              + f"  f32 {e['fp32_io']['value']/1e6:6.2f}  verified {d['verified']}", flush=True)
    except Exception as ex:
        print(spec, "FAILED", ex, p.stderr[-800:], flush=True)
