#!/usr/bin/env python
"""Memory timeline tool (counterpart of reference tools/plot_mem.py:8-297, which parses XLA buffer-assignment dumps).

There is no XLA here; the equivalent information comes from the CUDA caching allocator.  Record a snapshot around a
few training steps

    torch.cuda.memory._record_memory_history(max_entries=200000)
    ... train ...
    torch.cuda.memory._dump_snapshot("mem.pickle")

and run ``python tools/plot_mem.py mem.pickle [--top 20] [--png mem.png]`` to get the peak, the largest live
allocations at the peak (with the Python frame that made them) and, when matplotlib is available, a live-bytes plot.
"""
import argparse
import pickle
from collections import defaultdict


def load_events(path):
    snap = pickle.load(open(path, "rb"))
    events = []
    for dev_trace in snap.get("device_traces", []):
        for e in dev_trace:
            if e.get("action") in ("alloc", "free_completed", "free"):
                events.append(e)
    return events


def frame_of(e):
    for f in e.get("frames", []):
        fn = f.get("filename", "")
        if "site-packages/torch" not in fn and fn:
            return f"{fn.split('/')[-1]}:{f.get('line')} {f.get('name')}"
    return "?"


def analyse(events):
    live, cur, peak, peak_i = {}, 0, 0, 0
    series = []
    for i, e in enumerate(events):
        if e["action"] == "alloc":
            live[e["addr"]] = (e["size"], frame_of(e))
            cur += e["size"]
        elif e["addr"] in live:
            cur -= live.pop(e["addr"])[0]
        series.append(cur)
        if cur > peak:
            peak, peak_i, at_peak = cur, i, dict(live)
    return series, peak, peak_i, at_peak if peak else {}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("snapshot")
    p.add_argument("--top", type=int, default=20)
    p.add_argument("--png", default=None)
    a = p.parse_args()
    events = load_events(a.snapshot)
    series, peak, peak_i, at_peak = analyse(events)
    print(f"{len(events)} allocator events, peak live = {peak / 2**30:.3f} GiB at event {peak_i}")
    by_site = defaultdict(int)
    for size, site in at_peak.values():
        by_site[site] += size
    for site, size in sorted(by_site.items(), key=lambda x: -x[1])[:a.top]:
        print(f"  {size / 2**20:10.1f} MiB  {site}")
    if a.png:
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
            plt.figure(figsize=(10, 4))
            plt.plot([s / 2**30 for s in series])
            plt.axvline(peak_i, color="r", linestyle="--")
            plt.xlabel("allocator event")
            plt.ylabel("live GiB")
            plt.savefig(a.png, dpi=120, bbox_inches="tight")
            print("wrote", a.png)
        except ImportError:
            print("matplotlib not available; skipping the plot")


if __name__ == "__main__":
    main()
