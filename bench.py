#!/usr/bin/env python
"""bench.py -- audio-seconds per wall-second of the B200 path (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload vocoder|e2e]

A "step" is one pass of the hot path over one batch of synthetic input.  N=1 workload (default) is
BASELINE.json configs[1]: HiFi-GAN Generator only, batch 32 of 64-bin mel segments, 256 frames, hop 420,
11 413 Hz.  Under torchrun (N>1) every rank runs its own batch (utterance sharding, weak scaling), the
only collective being the final gather of the waveforms to rank 0, inside the timed region.

Rank 0 prints ONE JSON line.  `value` = whole-job audio-s/s with inputs resident in HBM; `e2e` = the same
metric through the public API (Generator.forward) with HOST buffers, H2D + D2H inside the timed region.
The default build is precision="fp16" (tcgen05 kind::f16 operands -- the same 10-bit mantissa as tf32 -- fp32
accumulate, fp32 residual streams; measured waveform SNR identical to the tf32 build); the tf32 build is timed in
the same run (`tf32_build`), and at N=1 a bounded end-to-end rtMRI -> wav sample goes into `pipeline`.
`--impl reference` times the CPU oracle port of the reference's Generator (the reference itself is Python
and cannot travel to the GPU box) on all host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

SR, HOP, NMELS = 11413, 420, 64
VOCODER_FLOP_PER_FRAME = 806.49e6   # 403.247 MMAC per mel frame, all 78 convs
ENGINE_FLOP_PER_FRAME = 2.0 * (403.247e6 - 0.094e6)  # everything but conv_post (32->1, CUDA-core kernel)


class AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def load_h():
    with open(os.path.join(ROOT, "config_custom.json"), "r", encoding="utf-8") as f:
        return AttrDict(json.load(f))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        rows = [ln for (t, ln) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [ln for (_, ln) in self.lines]
        for ln in rows:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def tf32_matmul_peak(device):
    """cuBLAS TF32 GEMM throughput on this GPU (context for the roofline; not the contract's `peak`)."""
    n = 8192
    a = torch.randn(n, n, device=device)
    b = torch.randn(n, n, device=device)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(2):
            a @ b
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return best


def cpu_reference_rate(batch, frames, repeats, threads=None, min_seconds=0.0):
    """The CPU arm: oracle port of models.Generator.forward (oracle/vocoder.py), fp32, all host cores."""
    from oracle.vocoder import generator_forward
    from mri2speech_b200.vocoder import Generator
    from mri2speech_b200 import synth
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    sd = {k: v.clone() for k, v in Generator(load_h()).state_dict().items()}
    mel = synth.synthetic_mels(batch, frames)
    with torch.no_grad():
        generator_forward(sd, load_h(), mel[:1, :, :32])  # warm-up
        times = []
        while len(times) < repeats or sum(times) < min_seconds:
            t0 = time.perf_counter()
            generator_forward(sd, load_h(), mel)
            times.append(time.perf_counter() - t0)
    audio_s = batch * frames * HOP / SR
    return audio_s / statistics.median(times), times, threads


def run_reference(args, rank, world):
    if rank != 0:
        return
    batch, frames = 1, 256  # bounded sample of config 2: one of the 32 segments per step (the reference
    # CLIs run B=1 per clip, scripts/run_mri_video_inference.py:241-242)
    rate, times, threads = cpu_reference_rate(batch, frames, max(args.steps, 1))
    ms = statistics.median(times) * 1e3
    sample = f"{batch} of 32 segments x {frames} frames per step ({batch * frames * HOP / SR:.1f} s audio)"
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": rate, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args), precision="fp32 on the host CPU (oracle port of models.Generator.forward)"),
        "cpu_baseline": {"value": rate, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def pipeline_sample(device, precision, n_clips=8):
    """Bounded end-to-end sample (BASELINE.json configs[2] shape, 8 instead of 64 clips): ragged uint8 clips of
    150-600 frames in pinned host memory -> H2D -> fused ingest -> encoder -> BiLSTM -> mel glue -> Generator -> D2H
    of the waveforms, through pipeline.MriToSpeech.infer; CUDA events around the whole call."""
    from mri2speech_b200 import synth
    from mri2speech_b200.acoustic import build_acoustic_model
    from mri2speech_b200.pipeline import MriToSpeech
    from mri2speech_b200.vocoder import Generator
    torch.manual_seed(1234)
    ac = build_acoustic_model(precision=precision)
    gen = Generator(load_h(), precision=precision)
    mean, std = synth.synthetic_scaler()
    pipe = MriToSpeech(ac, gen, mean, std, device)
    lens = synth.synthetic_lengths(64)[:n_clips]
    g = torch.Generator().manual_seed(1)
    clips = [torch.randint(0, 256, (ln, 256, 256), generator=g, dtype=torch.uint8).pin_memory() for ln in lens]
    frames = sum(lens)

    def run():
        out = pipe.infer(clips)
        return [o["audio"].to("cpu", non_blocking=True) for o in out]

    run()
    torch.cuda.synchronize()
    best = None
    for _ in range(2):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        wavs = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    assert all(w.numel() == ln * HOP for w, ln in zip(wavs, lens))
    audio_s = frames * HOP / SR
    return {"workload": f"rtMRI->wav, {n_clips} ragged uint8 clips ({frames} frames, {audio_s:.1f} s audio), host buffers",
            "ms": best, "value": audio_s / (best * 1e-3), "unit": "audio-s/s", "us_per_frame": best * 1e3 / frames,
            "h2d_bytes": frames * 256 * 256, "d2h_bytes": frames * HOP * 4, "precision": precision}


def workload_config(args):
    return {
        "workload": "BASELINE.json configs[1]: HiFi-GAN Generator only, batch 32 x 64-bin mel, 256 frames, "
                    "hop 420, 11413 Hz (per GPU)",
        "batch_per_gpu": args.batch, "frames": args.frames,
        "precision": ("fp16 operands (tcgen05 kind::f16; 10-bit mantissa like tf32), fp32 accumulate, fp32 residual / MRF "
                      "streams" if getattr(args, "precision", "fp16") == "fp16" else
                      "tf32 (tcgen05 kind::tf32, fp32 accumulate)"),
        "l2": "activations per stage (84-440 MB) exceed the 126 MB L2; no explicit flush",
        "sharding": "utterances per rank, final gather of waveforms to rank 0 inside the step (N>1)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default=os.environ.get("M2S_BENCH_PRECISION", "fp16"), choices=["tf32", "fp16"])
    ap.add_argument("--no-extras", action="store_true", help="skip the tf32-build and pipeline side measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: no CUDA device (there is no CPU fallback; use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=device)

    from mri2speech_b200 import _lib, synth
    from mri2speech_b200.vocoder import Generator

    torch.manual_seed(1234)
    gen = Generator(load_h(), precision=args.precision).to(device).eval()
    B, T = args.batch, args.frames
    mel_host = synth.synthetic_mels(B, T, seed=2024 + rank).pin_memory()
    mel = mel_host.to(device)
    audio_s_per_step = world * B * T * HOP / SR
    gather_buf = None
    if dist is not None and rank == 0:
        gather_buf = [torch.empty(B, 1, T * HOP, device=device) for _ in range(world)]

    def step():
        with torch.no_grad():
            wav = gen(mel)
        if dist is not None:
            dist.gather(wav, gather_buf, dst=0)
        return wav

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()

    # ---- timed region: device timing with CUDA events on the launching stream, max over ranks ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    _lib.profile(True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t1 = time.time()
    total_ms = e0.elapsed_time(e1)
    launch_ms, launch_flops = _lib.profile_read()
    _lib.profile(False)
    if dist is not None:
        tt = torch.tensor([total_ms], device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = audio_s_per_step / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the public API, H2D + D2H inside the timed region ----
    out_host = torch.empty(B, 1, T * HOP).pin_memory()

    def e2e_step():
        x = mel_host.to(device, non_blocking=True)
        with torch.no_grad():
            wav = gen(x)
        out_host.copy_(wav, non_blocking=True)

    for _ in range(3):
        e2e_step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e2, e3 = torch.cuda.Event(True), torch.cuda.Event(True)
    e2.record()
    for _ in range(args.steps):
        e2e_step()
    e3.record()
    torch.cuda.synchronize()
    e2e_ms = e2.elapsed_time(e3)
    if dist is not None:
        tt = torch.tensor([e2e_ms], device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    e2e_value = audio_s_per_step / (e2e_ms / args.steps * 1e-3)

    # ---- side measurements (rank 0, N=1): the tf32 build on the same workload; a bounded rtMRI -> wav sample ----
    tf32_build = pipeline_info = graph_info = None
    if world == 1 and not args.no_extras:
        # the same forward replayed as ONE CUDA graph (mri2speech_b200/graphs.py): no per-launch host cost, no
        # per-launch profiling events -- what a fixed-shape serving loop gets
        from mri2speech_b200.graphs import graph_generator
        gg = graph_generator(gen, mel)
        for _ in range(3):
            gg(mel)
        torch.cuda.synchronize()
        ea, eb = torch.cuda.Event(True), torch.cuda.Event(True)
        ea.record()
        for _ in range(args.steps):
            gg(mel)
        eb.record()
        torch.cuda.synchronize()
        msg = ea.elapsed_time(eb) / args.steps
        graph_info = {"ms_per_step": msg, "value": audio_s_per_step / (msg * 1e-3), "unit": "audio-s/s",
                      "steps": args.steps, "note": "CUDA-graph replay of Generator.forward, input copied into the static buffer each step"}
        del gg
        if args.precision != "tf32":
            torch.manual_seed(1234)
            gen32 = Generator(load_h(), precision="tf32").to(device).eval()
            with torch.no_grad():
                for _ in range(3):
                    gen32(mel)
                torch.cuda.synchronize()
                k = max(3, min(args.steps, 5))
                ea, eb = torch.cuda.Event(True), torch.cuda.Event(True)
                ea.record()
                for _ in range(k):
                    gen32(mel)
                eb.record()
                torch.cuda.synchronize()
            ms32 = ea.elapsed_time(eb) / k
            tf32_build = {"ms_per_step": ms32, "value": audio_s_per_step / (ms32 * 1e-3), "unit": "audio-s/s", "steps": k}
            del gen32
        pipeline_info = pipeline_sample(device, args.precision)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (conv_engine_kernel, tensor-pipe bound) ----
    peaks, peak_src = measured_peaks()
    n_launch = len(launch_ms)
    per_step_launches = n_launch // max(args.steps, 1)
    engine_ms_per_step = sum(launch_ms) / max(args.steps, 1)
    alg_flop_per_launch = ENGINE_FLOP_PER_FRAME * B * T / max(per_step_launches, 1)
    avg_launch_ms = engine_ms_per_step / max(per_step_launches, 1)
    achieved = alg_flop_per_launch / (avg_launch_ms * 1e-3) / 1e12 if avg_launch_ms > 0 else 0.0
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    tf32_peak = tf32_matmul_peak(device)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_conv_engine.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    roofline = {
        "kernel": f"conv_engine_kernel / conv_engine_pair_kernel (tcgen05 kind::{'f16' if args.precision == 'fp16' else 'tf32'} "
                  "implicit-GEMM conv, all 77 GEMM-shaped layers)",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic,
        "peak_source": f"{peak_src}: dense bf16 cuBLAS, sustained (kind::f16 issues at the bf16 rate, kind::tf32 at half)",
        "tf32_cublas_tflops_measured_here": tf32_peak, "frac_of_tf32_cublas": achieved / tf32_peak if tf32_peak else None,
        "launches_per_step": per_step_launches, "engine_ms_per_step": engine_ms_per_step,
        "engine_share_of_step": engine_ms_per_step / ms_per_step,
        "executed_tflop_per_step": sum(launch_flops) / max(args.steps, 1) / 1e12,
        "algorithmic_tflop_per_step": ENGINE_FLOP_PER_FRAME * B * T / 1e12,
        # mean duration of each of the engine launches of one forward, in launch order (conv_pre, then per stage:
        # ups, 3 ResBlocks x 3 x (conv1, conv2)), microseconds
        "per_launch_us": [round(1e3 * sum(launch_ms[i::per_step_launches]) / max(args.steps, 1), 1)
                          for i in range(per_step_launches)] if per_step_launches else [],
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        rate, times, threads = cpu_reference_rate(1, T, 3, min_seconds=10.0)
        cpu = {"value": rate, "unit": "audio-s/s", "cores": threads, "kind": "port",
               "sample": f"1 of {B} segments x {T} frames (B=1 as the reference CLI runs it), median of {len(times)} "
                         f"passes ({sum(times):.1f} s CPU wall)"}

    line = {
        "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": workload_config(args), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": mel_host.numel() * 4 * world,
                "d2h_bytes_per_step": out_host.numel() * 4 * world, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(gen.launches_per_forward()) * args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "tf32_build": tf32_build, "graph_replay": graph_info,
        "pipeline": pipeline_info,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
