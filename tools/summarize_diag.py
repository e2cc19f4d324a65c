"""Print compact summaries of gpurun_out/diag_*.json (helper for bring-up runs)."""
import glob
import json
import os
import sys

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def load(name):
    path = os.path.join(OUT, f"diag_{name}.json")
    return json.load(open(path)).get("results", []) if os.path.exists(path) else []


def main():
    args = sys.argv[1:]
    if len(args) >= 3 and args[0] == "compare":
        a = {r["name"]: r for r in load(args[1])}
        b = {r["name"]: r for r in load(args[2])}
        for k in a:
            if k in b and "ms" in a[k]:
                print(f"{k:30s} {args[1]} {a[k]['ms'] * 1e3:8.1f} us   {args[2]} {b[k]['ms'] * 1e3:8.1f} us   "
                      f"{b[k]['tflops']:7.1f} TF/s")
        return
    for name in args:
        res = load(name)
        bad = [r for r in res if r.get("tc_nan") or r.get("tc_err32", 0) > 1e-4 or r.get("tc_vs_simt32", 0) > 1e-4
               or r.get("tc_err16", 0) > 0.02]
        print(name, "cases", len(res), "bad", bad)
        for r in res:
            if "ms" in r:
                print(f"   {r['name']:30s} {r['ms'] * 1e3:8.1f} us {r['tflops']:7.1f} TF/s")


if __name__ == "__main__":
    main()
