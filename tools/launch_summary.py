"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of one
training step (the launches between two consecutive im2col kernels = one rpo_forward+rpo_backward)."""
import csv, re, sys, collections

def main(path, out=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        if unit in ("us", "usecond"): v *= 1e3
        elif unit in ("ms", "msecond"): v *= 1e6
        name = r["Kernel Name"]
        rows.append((name, v))
    starts = [i for i, (n, _) in enumerate(rows) if "im2col" in n]
    if len(starts) < 2:
        print("need two steps in the capture"); return
    a, b = starts[-2], starts[-1]
    step = rows[a:b]
    agg = collections.OrderedDict()
    for n, v in step:
        short = re.sub(r"\(.*", "", n)
        short = re.sub(r"^void ", "", short)
        short = short.replace("rpo::", "")
        d = agg.setdefault(short, [0, 0.0])
        d[0] += 1; d[1] += v
    total = sum(v for _, v in step)
    lines = [f"one step = {len(step)} launches, {total/1e6:.3f} ms summed kernel time (ncu, serialised, cold cache)"]
    lines.append(f"{'kernel':90s} {'launches':>8s} {'total us':>10s} {'avg us':>9s} {'share':>7s}")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k[:90]:90s} {c:8d} {v/1e3:10.1f} {v/1e3/c:9.2f} {100*v/total:6.1f}%")
    text = "\n".join(lines)
    print(text)
    if out:
        open(out, "w").write(text + "\n")

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
