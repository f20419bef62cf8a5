"""Summarise gpurun_out/conv_launch_table.json (bench.py --dump-kernels): in-step time per conv shape."""
import collections
import json
import sys

t = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv_launch_table.json"))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in t:
    a = agg[(r["kernel"], tuple(r["n_h_w_cin_cout_k_cin2"][1:]))]
    a[0] += 1; a[1] += r["us"]; a[2] += r["tflops"] * r["us"]
tot = collections.Counter()
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
print(f"{'kernel':11s} {'h,w,cin,cout,k,cin2':27s}  n  total_us  avg_us TFLOP/s")
for i, (k, (n, us, fl)) in enumerate(sorted(agg.items(), key=lambda kv: -kv[1][1])):
    if i < top:
        print(f"{k[0]:11s} {str(k[1]):27s} {n:2d} {us:9.1f} {us / n:7.1f} {fl / us:7.1f}")
    tot[k[0]] += us
print(dict(tot))
