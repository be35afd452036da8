"""Per-launch table of the LAST pass in an ncu launch list: python tools/launch_table.py FILE.csv [first-kernel-marker] [from-index]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import summarize_launches as s

ls = s.load(sys.argv[1])
marker = sys.argv[2] if len(sys.argv) > 2 else "stem"
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
starts = [i for i, x in enumerate(ls) if marker in x["name"]]
ls = ls[starts[-1]:]
tot = 0.0
agg = {}
for i, x in enumerate(ls):
    n = x["name"]
    nm = "mb_expand_dw" if "mb_expand" in n else "mb_project" if "mb_project" in n else "fused_er" if "fused_er" in n else s.short(n)
    t = x["gpu__time_duration.sum"] / 1000
    tot += t
    a = agg.setdefault(nm, [0, 0.0])
    a[0] += 1; a[1] += t
    if i >= lo:
        print(i, nm, f"{t:7.1f}us rd {x.get('dram__bytes_read.sum',0)/1e6:7.1f} wr {x.get('dram__bytes_write.sum',0)/1e6:7.1f}MB "
              f"tensor {x.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',0):.1f}% issue {x.get('smsp__issue_active.avg.pct_of_peak_sustained_active',0):.1f}%")
print(f"total {tot:.1f} us over {len(ls)} launches")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} {c:3d} {t:8.1f} us {100*t/tot:5.1f}%")
