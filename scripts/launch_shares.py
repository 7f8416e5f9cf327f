"""Per-kernel share of a step from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X ...).
usage: python scripts/launch_shares.py <launches.csv> <out.json> "<command that produced it>" """
import csv
import json
import re
import sys


def main():
    src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    agg = {}
    for r in rows[1:]:
        if len(r) <= iv or "gpu__time_duration" not in ",".join(r):
            continue
        name = re.sub(r"^void ", "", r[ik].split("(")[0].strip()).replace("(bool)", "")
        us = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    out = dict(command=cmd, note="per-launch times under ncu are cold-cache and serialised: use the SHARES",
               kernels={k: dict(launches=a[0], total_ms=a[1] / 1e3, share=a[1] / tot, avg_us=a[1] / a[0])
                        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])})
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: round(v["share"], 4) for k, v in out["kernels"].items()}))


if __name__ == "__main__":
    main()
