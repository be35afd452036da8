"""profiles/traffic_encoder_engine.json from an encoder launch list (ncu --csv with dram__bytes_read/write, one pass):

    python tools/make_traffic_json.py profiles/launches_r02_encoder_final.csv "<how the list was taken>"

bench.py reads it for `roofline.traffic` (DRAM bytes per launch of the encoder's tcgen05 kernels, the same launches its
live CUDA-event leg averages)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import summarize_launches as s

TC = ("conv_engine_pair_kernel", "conv_engine_kernel", "fused_er_kernel", "mb_expand_dw_kernel", "mb_project_kernel")


def main():
    path, how = sys.argv[1], sys.argv[2]
    ls = s.load(path)
    starts = [i for i, x in enumerate(ls) if "stem" in x["name"]]
    ls = ls[starts[-1]:]
    per, tot_t, tot_b, n, tp = {}, 0.0, 0.0, 0, 0.0
    for x in ls:
        key = next((k for k in TC if k in x["name"]), None)
        if key is None:
            continue
        t = x["gpu__time_duration.sum"] / 1000.0
        b = x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0)
        a = per.setdefault(key, {"launches": 0, "us": 0.0, "dram_MB": 0.0, "tensor_pipe_pct_time_weighted": 0.0})
        a["launches"] += 1; a["us"] += t; a["dram_MB"] += b / 1e6
        a["tensor_pipe_pct_time_weighted"] += x.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * t
        tot_t += t; tot_b += b; n += 1
        tp += x.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0) * t
    for a in per.values():
        a["tensor_pipe_pct_time_weighted"] /= max(a["us"], 1e-9)
        a["dram_GBps"] = a["dram_MB"] / max(a["us"], 1e-9) * 1e3
    out = {"source": f"{os.path.basename(path)}: {how}", "frames": int(sys.argv[3]) if len(sys.argv) > 3 else 2048, "tc_launches": n, "tc_us": tot_t,
           "dram_bytes_per_pass": tot_b, "dram_bytes_per_launch": tot_b / max(n, 1),
           "tensor_pipe_active_pct_time_weighted": tp / max(tot_t, 1e-9), "per_kernel": per}
    dst = os.path.join(os.path.dirname(path), "traffic_encoder_engine.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
