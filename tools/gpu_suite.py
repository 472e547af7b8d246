"""Bring-up driver for the GPU box: runs each GPU test function in its own process (a CUDA fault in one cannot
poison the others), with a timeout, and writes logs + a summary under gpurun_out/.
    python tools/gpu_suite.py [substring filters...]"""
import json
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "gpurun_out")
os.makedirs(OUT, exist_ok=True)


def collect():
    r = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "--collect-only", "-q"], cwd=REPO,
                       capture_output=True, text=True)
    fns = []
    for line in r.stdout.splitlines():
        if "::" in line:
            fn = line.split("[")[0]
            if fn not in fns:
                fns.append(fn)
    return fns


def main():
    filters = sys.argv[1:]
    fns = [f for f in collect() if not filters or any(s in f for s in filters)]
    summary = {}
    t_all = time.time()
    for fn in fns:
        name = fn.replace("/", "_").replace("::", "-").replace(".py", "")
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", fn, "-q", "-x", "--no-header", "-p", "no:cacheprovider"],
                               cwd=REPO, capture_output=True, text=True, timeout=420)
            out, rc = r.stdout + r.stderr, r.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = (e.stdout or b"").decode(errors="replace") + "\nTIMEOUT", -9
        open(os.path.join(OUT, name + ".log"), "w").write(out)
        tail = [l for l in out.splitlines() if l.strip()][-1:] or [""]
        summary[fn] = {"rc": rc, "sec": round(time.time() - t0, 1), "tail": tail[0][:200]}
        print("%-90s rc=%d %.0fs %s" % (fn, rc, time.time() - t0, tail[0][:120]), flush=True)
        if rc != 0:
            lines = out.splitlines()
            errs = [l for l in lines if l.startswith("E ") or "Error" in l or "s2e:" in l][:12]
            print("    " + "\n    ".join(errs), flush=True)
    json.dump(summary, open(os.path.join(OUT, "suite_summary.json"), "w"), indent=1)
    print("total %.0fs, failed: %d / %d" % (time.time() - t_all, sum(1 for v in summary.values() if v["rc"] != 0), len(summary)))


if __name__ == "__main__":
    main()
