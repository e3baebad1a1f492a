#!/usr/bin/env python
"""Run a measurement command with the nvidia-smi clock sampler around it and keep both in one file:

    python tools/with_clocks.py profiles/r02_bn_kernels_gbs.json -- python tools/bench_bn.py

The command's JSON lines become `records`; `clocks` = median SM clock under load, max clock, throttle reasons seen
during the run (B200_PROFILING.md: a number kept under profiles/ carries its clock record)."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.clocks import ClockSampler  # noqa: E402


def main():
    out, sep = sys.argv[1], sys.argv.index("--")
    cmd = sys.argv[sep + 1:]
    s = ClockSampler(int(os.environ.get("CLOCK_GPU", "0")))
    s.start()
    r = subprocess.run(cmd, capture_output=True, text=True)
    clocks = s.stop()
    recs, other = [], []
    for line in r.stdout.splitlines():
        line = line.strip()
        if line.startswith("{"):
            try:
                recs.append(json.loads(line))
                continue
            except ValueError:
                pass
        if line:
            other.append(line)
    doc = {"command": " ".join(cmd), "returncode": r.returncode, "clocks": clocks, "records": recs}
    if other:
        doc["stdout"] = other[-20:]
    if r.returncode != 0:
        doc["stderr_tail"] = r.stderr.splitlines()[-15:]
    with open(out, "w") as fh:
        json.dump(doc, fh, indent=1)
    print(out, "rc", r.returncode, "records", len(recs), "clocks", clocks)


if __name__ == "__main__":
    main()
