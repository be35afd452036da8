"""Summarise an `ncu --csv` launch list (gpu__time_duration.sum [+ dram bytes, tensor-pipe %]) per kernel.

    python tools/summarize_launches.py FILE.csv [marker-substring-of-first-kernel-of-a-pass]
"""
import collections
import csv
import json
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hdr]
    ki, mi, mn, idi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("ID")
    out = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= mi:
            continue
        d = out.setdefault(r[idi], {"name": r[ki]})
        d[r[mn]] = float(r[mi].replace(",", ""))
    return list(out.values())


def short(name):
    for key in ("conv_engine_pair_kernel", "conv_engine_kernel", "resblock_pair_kernel", "dwconv_tma_kernel",
                "dwconv_kernel", "se_scale_kernel", "se_mlp_kernel", "se_apply_kernel", "im2col_s2_kernel",
                "stem_rows_kernel", "stem_kernel", "gap_kernel", "zero_rows_kernel", "conv_post_kernel",
                "bct_to_btc_kernel", "frame_minmax_kernel", "lstm_recurrence_kernel"):
        if key in name:
            return key
    return name[:40]


def main():
    ls = load(sys.argv[1])
    marker = sys.argv[2] if len(sys.argv) > 2 else None
    if marker:
        starts = [i for i, x in enumerate(ls) if marker in x["name"]]
        ls = ls[starts[-2]:starts[-1]] if len(starts) >= 2 else ls[starts[-1]:]
    t = collections.Counter()
    c = collections.Counter()
    b = collections.Counter()
    tp = collections.Counter()
    for x in ls:
        k = short(x["name"])
        dur = x["gpu__time_duration.sum"]
        t[k] += dur
        c[k] += 1
        b[k] += x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0)
        tp[k] += x.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0) * dur
    tot = sum(t.values())
    print(f"{len(ls)} launches, {tot / 1e3:.1f} us")
    res = {}
    for k, v in t.most_common():
        res[k] = {"launches": c[k], "us": v / 1e3, "share": v / tot, "dram_MB": b[k] / 1e6,
                  "dram_GBps": b[k] / v if v else 0, "tensor_pipe_pct": tp[k] / v if v else 0}
        print(f"{k:28s} {c[k]:4d} {v / 1e3:9.1f} us {100 * v / tot:5.1f}%  {b[k] / 1e6:9.1f} MB  {b[k] / v if v else 0:7.0f} GB/s  "
              f"tensor {tp[k] / v if v else 0:5.1f}%")
    return res


if __name__ == "__main__":
    main()
