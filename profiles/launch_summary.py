#!/usr/bin/env python3
"""ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X`) -> per-kernel table of ONE hybrid step.

    python profiles/launch_summary.py gpurun_out/launches_step.csv "comment line" > profiles/rNN_launches.csv

A step starts at cvt_queries_kernel (first kernel of rr_dense_topk); the second step found in the list is used
(the first one is the warm-up).  Times are cold-cache and serialised: compare SHARES with bench.py's live numbers."""
import csv
import sys


def main(path, note):
    rows = []
    for r in csv.reader(l for l in open(path) if l.startswith('"')):
        if r[0] == "ID":
            hdr = r
            continue
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = d["Kernel Name"].split("(")[0].split("::")[-1].replace("void ", "")
        unit = d["Metric Unit"]
        v = float(d["Metric Value"].replace(",", ""))
        us = v / 1000.0 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1000.0
        rows.append((name, us))
    starts = [i for i, (n, _) in enumerate(rows) if n.startswith("cvt_queries_kernel")]
    if len(starts) < 2:
        raise SystemExit("need at least two steps in the launch list")
    a = starts[1]
    b = starts[2] if len(starts) > 2 else len(rows)
    step = [(n, t) for n, t in rows[a:b] if not n.startswith(("at::", "vectorized", "elementwise", "unrolled", "reduce_kernel"))]
    # a step ends at the fusion kernel
    for i, (n, _) in enumerate(step):
        if n.startswith("fuse_topk_kernel"):
            step = step[:i + 1]
            break
    total = sum(t for _, t in step)
    agg = {}
    for n, t in step:
        c = agg.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += t
    print(f"# {note}")
    print("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event numbers, not absolutes")
    print("# kernel, launches, total_us, share_of_step")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n},{c},{t:.1f},{t / total:.4f}")
    print()
    print("# launch sequence of the step (kernel, us)")
    for n, t in step:
        print(f"{n},{t:.1f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu launch list")
