"""Interleaved A/B of environment switches on the headline bench (ABAB... in separate processes, medians reported).

Back-to-back benches on one box drift by 3-4 % as the GPU warms up into its power cap (profiles/README.md, last-shot
run), so a single A-then-B comparison cannot resolve a 3 % change.  Usage (on the GPU box):

    python tools/ab_bench.py --b WXF_PDL=1 --rounds 3
    python tools/ab_bench.py --a WXF_ATTN_V2=0 --b WXF_ATTN_V2=1 --rounds 2 --steps 5
"""
import argparse
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


EXTRA = []


def run(env_kv, steps):
    env = dict(os.environ)
    for kv in env_kv:
        k, v = kv.split("=", 1)
        env[k] = v
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(steps), "--warmup", "3",
                          "--no-cpu-baseline"] + EXTRA, env=env, capture_output=True, text=True, timeout=300)
    for ln in reversed(out.stdout.strip().splitlines()):
        try:
            return json.loads(ln)
        except ValueError:
            continue
    raise RuntimeError(out.stderr[-2000:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--a", nargs="*", default=[], help="KEY=VALUE switches of arm A (default: none)")
    ap.add_argument("--b", nargs="*", default=[], help="KEY=VALUE switches of arm B")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--workload", default=None, help="bench workload (default: the headline 0.25 deg one)")
    args = ap.parse_args()
    if args.workload:
        EXTRA.extend(["--workload", args.workload])
    res = {"A": [], "B": []}
    fam = {"A": {}, "B": {}}
    for r in range(args.rounds):
        for arm, kv in (("A", args.a), ("B", args.b)):
            d = run(kv, args.steps)
            res[arm].append(d["ms_per_step"])
            for k, v in d["kernel_families"].items():
                fam[arm].setdefault(k, []).append(v["ms"])
            print(f"round {r} {arm} {' '.join(kv) or '(default)'}: {d['ms_per_step']:.3f} ms/step, e2e "
                  f"{d['e2e']['value']:.2f} steps/s, clocks {d['clocks']}", flush=True)
    ma, mb = statistics.median(res["A"]), statistics.median(res["B"])
    print(f"median ms/step: A {ma:.3f}  B {mb:.3f}  (B/A = {mb / ma:.4f})")
    for k in fam["A"]:
        a, b = statistics.median(fam["A"][k]), statistics.median(fam["B"].get(k, [float('nan')]))
        print(f"  {k:16s} A {a:7.3f} ms   B {b:7.3f} ms")


if __name__ == "__main__":
    main()
