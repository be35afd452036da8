"""profiles/traffic_conv_engine.json from a vocoder launch list (ncu --csv of `bench.py --workload vocoder`): the LAST
forward (from conv_pre's glue / layout kernel to conv_post) of the list.

    python tools/make_vocoder_traffic_json.py profiles/launches_r02_vocoder_final.csv "<how the list was taken>"
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import summarize_launches as s

ENGINE = ("conv_engine_pair_kernel", "conv_engine_kernel", "resblock_pair_kernel")


def main():
    path, how = sys.argv[1], sys.argv[2]
    ls = s.load(path)
    ends = [i for i, x in enumerate(ls) if "conv_post_kernel" in x["name"]]
    lo = ends[-2] + 1 if len(ends) >= 2 else 0
    ls = ls[lo:ends[-1] + 1]
    per, tot_t, tot_b, n, tp, all_t = {}, 0.0, 0.0, 0, 0.0, 0.0
    for x in ls:
        t = x["gpu__time_duration.sum"] / 1000.0
        all_t += t
        key = next((k for k in ENGINE if k in x["name"]), None)
        if key is None:
            continue
        b = x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0)
        pipe = x.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
        a = per.setdefault(key, {"launches": 0, "us": 0.0, "dram_MB": 0.0, "tensor_pipe_pct": 0.0})
        a["launches"] += 1; a["us"] += t; a["dram_MB"] += b / 1e6; a["tensor_pipe_pct"] += pipe * t
        tot_t += t; tot_b += b; n += 1; tp += pipe * t
    for a in per.values():
        a["tensor_pipe_pct"] /= max(a["us"], 1e-9)
        a["share"] = a["us"] / max(all_t, 1e-9)
        a["dram_GBps"] = a["dram_MB"] / max(a["us"], 1e-9) * 1e3
    out = {"source": f"{os.path.basename(path)}: {how}", "engine_launches": n, "forward_us": all_t,
           "dram_bytes_per_forward": tot_b, "dram_bytes_per_launch": tot_b / max(n, 1),
           "tensor_pipe_active_pct_time_weighted": tp / max(tot_t, 1e-9), "per_kernel": per}
    dst = os.path.join(os.path.dirname(path), "traffic_conv_engine.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
