#!/usr/bin/env python
"""bench.py -- audio-seconds per wall-second, rtMRI -> wav, of the B200 path (contract: DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|e2e|sharded|vocoder]

A "step" is one pass of the hot path over one batch of synthetic input.

  * N = 1 (default; `--workload e2e`): BASELINE.json configs[2] -- end-to-end rtMRI -> mel -> wav on 64 synthetic clips of
    150-600 frames (lengths from seed 4321, ~24 k frames, 1.5 GB of raw uint8 frames).  One step = the whole batch through
    `pipeline.MriToSpeech.infer` (fused uint8 ingest -> frame-CNN encoder -> BiLSTM -> mel glue -> HiFi-GAN Generator).
      value : device-timed (CUDA events), clips already resident in HBM, waveforms left in HBM;
      e2e   : the same call with the clips in PINNED HOST memory and the waveforms returned in pinned host memory --
              H2D and D2H inside the timed region (overlapped with compute by infer's copy streams).
  * N > 1 under torchrun (`--workload sharded`): BASELINE.json configs[3] -- 4 096 clips sharded by utterance.  The job is
    cut into 8 waves of 512 clips; one step = one wave: longest-processing-time shards over the N ranks, every rank
    runs its shard, the waveforms are gathered to rank 0 inside the step (the path's only collective).  Total work per
    step does not depend on N: "scaling": "strong".
  * `--workload vocoder`: BASELINE.json configs[1] (HiFi-GAN Generator only, 32 x 256 mel frames); at N = 1 it is also
    measured as the side key `vocoder` with its own roofline (the tensor-pipe target of the metric is on these convs).

Rank 0 prints ONE JSON line.  `--impl reference` times the CPU oracle port of the same chain (restated encoder +
torch.nn.LSTM + restated Generator; the reference itself is Python + un-vendored timm and cannot travel to the GPU box)
on all host cores, one clip per step (B = 1, as scripts/run_mri_video_inference.py runs), on a bounded sample.
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
# algorithmic work per frame: 2 * MAC
ENC_GEMM_FLOP = 2.0 * 1485.6e6        # dense convs of the encoder (tcgen05 engine)
ENC_FLOP = 3.011e9                    # + depthwise / stem
RNN_FLOP = 8.77e6
VOC_FLOP = 806.49e6
VOC_ENGINE_FLOP = 2.0 * (403.247e6 - 0.094e6)   # everything but conv_post (32 -> 1, CUDA-core kernel)
PATH_FLOP = 3.826e9
DEFAULT_PRECISION = "fp16"
N_BASE, BASE_FRAMES = 16, 600         # unique synthetic clips behind the clip lists (compute does not depend on the data)
WAVE_CLIPS, JOB_CLIPS = 512, 4096


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
            return json.load(f), "measured (MEASURED_PEAKS.json)"
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


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
def base_clips():
    """N_BASE unique raw uint8 clips of BASE_FRAMES frames (smooth synthetic fields)."""
    from mri2speech_b200 import synth
    return [synth.synthetic_clip_u8(1000 + i, BASE_FRAMES) for i in range(N_BASE)]


def clip_views(base, lengths, ids):
    return [base[i % len(base)][: lengths[i]] for i in ids]


def build_pipeline(device, precision, max_batch_frames):
    from mri2speech_b200 import synth
    from mri2speech_b200.acoustic import build_acoustic_model
    from mri2speech_b200.pipeline import MriToSpeech
    from mri2speech_b200.vocoder import Generator
    torch.manual_seed(1234)
    ac = build_acoustic_model(precision=precision)
    gen = Generator(load_h(), precision=precision)
    mean, std = synth.synthetic_scaler()
    pipe = MriToSpeech(ac, gen, mean, std, device)
    pipe.reserve(max_batch_frames, BASE_FRAMES)
    return pipe


def workload_config(args, world):
    prec = ("fp16 operands (tcgen05 kind::f16; 10-bit mantissa like tf32), fp32 accumulate, fp16 residual streams, fp32 MRF sums"
            if args.precision == "fp16" else "tf32 (tcgen05 kind::tf32, fp32 accumulate)")
    if args.workload == "vocoder":
        wl = ("BASELINE.json configs[1]: HiFi-GAN Generator only, batch 32 x 64-bin mel, 256 frames, hop 420, 11413 Hz "
              "(per GPU)")
    elif args.workload == "sharded":
        wl = (f"BASELINE.json configs[3]: utterance-sharded rtMRI->wav, {JOB_CLIPS} synthetic clips of 150-600 frames "
              f"(seed 4321) in {JOB_CLIPS // WAVE_CLIPS} waves of {WAVE_CLIPS}; one step = one wave, LPT shards over the "
              "ranks, final gather of the waveforms to rank 0 inside the step")
    else:
        wl = ("BASELINE.json configs[2]: end-to-end rtMRI->mel->wav, batch 64 synthetic clips of 150-600 frames "
              "(variable length, seed 4321), raw uint8 256x256 frames, 1 x B200")
    return {
        "workload": wl, "precision": prec, "max_batch_frames": args.max_batch_frames,
        "l2": "one step streams 1.66 GB (N=1) of frames and multi-GB activations: far beyond the 126 MB L2, no explicit flush",
        "clips": f"{N_BASE} unique synthetic uint8 clips of {BASE_FRAMES} frames behind the clip list (clip i = "
                 "base[i % 16][:len_i]); kernels are data-independent",
        "sharding": "utterances per rank (LPT on frames), no collective on the hot path, one gather per step" if world > 1
                    else "single GPU",
    }


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the whole chain, one clip per step
# ---------------------------------------------------------------------------------------------------------------------
class CpuChain:
    def __init__(self, threads=None):
        from mri2speech_b200 import synth
        from mri2speech_b200.acoustic import build_acoustic_model
        from mri2speech_b200.vocoder import Generator
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        torch.manual_seed(1234)
        self.sd_ac = {k: v.clone() for k, v in build_acoustic_model().state_dict().items()}
        self.h = load_h()
        self.sd_gen = {k: v.clone() for k, v in Generator(self.h).state_dict().items()}
        self.mean, self.std = synth.synthetic_scaler()
        self.synth = synth

    def clip(self, frames, clip_id=0):
        return self.synth.synthetic_clip(clip_id, frames)

    def run(self, clip):
        """scripts/run_mri_video_inference.py:218-243 on the host: frames -> normalised mel -> log-mel -> waveform."""
        from oracle.acoustic import acoustic_forward
        from oracle.glue import mel_glue
        from oracle.vocoder import generator_forward
        t0 = time.perf_counter()
        with torch.no_grad():
            mel = acoustic_forward(self.sd_ac, clip[None, :, None])[0]
            _, _, voc_in = mel_glue(mel, self.mean, self.std)
            wav = generator_forward(self.sd_gen, self.h, voc_in.unsqueeze(0))
        assert wav.shape[-1] == clip.shape[0] * HOP
        return time.perf_counter() - t0


def run_reference(args, rank, world):
    if rank != 0:
        return
    chain = CpuChain()
    probe = chain.clip(8)
    chain.run(probe)                       # first-touch / thread-pool warm-up
    per_frame = chain.run(probe) / 8.0
    # bounded sample: one clip per step, sized so that warmup + steps end in ~150 s; configs[0]'s 150 frames if they fit
    budget_s = 150.0
    frames = int(max(16, min(150, budget_s / max(args.steps + args.warmup, 1) / max(per_frame, 1e-6))))
    clip = chain.clip(frames)
    for _ in range(args.warmup):
        chain.run(clip)
    times = [chain.run(clip) for _ in range(max(args.steps, 1))]
    audio_s = frames * HOP / SR
    total = sum(times)
    rate = audio_s * len(times) / total
    sample = (f"one synthetic clip of {frames} frames per step ({audio_s:.2f} s audio; B=1 as scripts/run_mri_video_inference.py "
              f"runs), {len(times)} steps, {total:.1f} s CPU wall")
    args.workload = "sharded" if world > 1 else "e2e"
    line = {
        "impl": "reference", "metric": "audio_seconds_per_second", "value": rate, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": dict(workload_config(args, world), precision="fp32 on the host CPU: oracle port of the chain "
                       "(restated tf_efficientnetv2_b2 encoder + torch.nn.LSTM + restated models.Generator)"),
        "cpu_baseline": {"value": rate, "unit": "audio-s/s", "cores": chain.threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# timing helpers
# ---------------------------------------------------------------------------------------------------------------------
def timed(step, steps, dist=None, device=None):
    """K steps bracketed by barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t0 = time.time()
    e0.record()
    for k in range(steps):
        step(k)
    e1.record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        tt = torch.tensor([ms], device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    return ms, t0, t1


def breakdown(ms, flops, tags, steps):
    from mri2speech_b200._lib import PROFILE_TAGS
    out = {}
    for name in PROFILE_TAGS.values():
        out[name] = {"ms_per_step": 0.0, "launch_groups_per_step": 0.0, "executed_tflop_per_step": 0.0}
    for m, f, t in zip(ms, flops, tags):
        b = out[PROFILE_TAGS.get(t, "other")]
        b["ms_per_step"] += m / steps
        b["launch_groups_per_step"] += 1.0 / steps
        b["executed_tflop_per_step"] += f / steps / 1e12
    return {k: v for k, v in out.items() if v["launch_groups_per_step"] > 0}


def vocoder_side(device, precision, steps, peaks):
    """BASELINE.json configs[1] on this GPU: Generator only, 32 x 256 mel frames, device-resident; engine roofline."""
    from mri2speech_b200 import _lib, synth
    from mri2speech_b200.vocoder import Generator
    torch.manual_seed(1234)
    gen = Generator(load_h(), precision=precision).to(device).eval()
    B, T = 32, 256
    mel = synth.synthetic_mels(B, T).to(device)
    with torch.no_grad():
        for _ in range(3):
            gen(mel)
        torch.cuda.synchronize()
        _lib.profile(True)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(steps):
            gen(mel)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lms, lfl, ltg = _lib.profile_read(with_tags=True)
    _lib.profile(False)
    eng = sum(m for m, t in zip(lms, ltg) if t == 4) / steps
    n_eng = sum(1 for t in ltg if t == 4) // steps
    achieved = VOC_ENGINE_FLOP * B * T / (eng * 1e-3) / 1e12 if eng > 0 else 0.0
    burst = float(peaks.get("bf16_tflops", 0.0))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_conv_engine.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    audio_s = B * T * HOP / SR
    return {"workload": "BASELINE.json configs[1]: Generator only, 32 x 256 mel frames, device-resident", "precision": precision,
            "ms_per_step": ms, "value": audio_s / (ms * 1e-3), "unit": "audio-s/s", "steps": steps,
            "roofline": {"kernel": "conv_engine_kernel / conv_engine_pair_kernel / resblock_pair_kernel (all vocoder GEMM launches)",
                         "bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s",
                         "frac": achieved / burst if burst else None,
                         "peak_source": "MEASURED_PEAKS.json bf16 BURST (timed region < 1 s)",
                         "engine_ms_per_step": eng, "engine_launches_per_step": n_eng,
                         "traffic_capture": traffic}}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "e2e", "sharded", "vocoder"])
    ap.add_argument("--precision", default=os.environ.get("M2S_BENCH_PRECISION", DEFAULT_PRECISION), choices=["tf32", "fp16"])
    ap.add_argument("--max-batch-frames", type=int, default=int(os.environ.get("M2S_BENCH_MBF", "8192")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the side measurements (other build, vocoder-only)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "auto":
        args.workload = "sharded" if world > 1 else "e2e"

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
            os.environ["NCCL_DEBUG"] = "WARN"
        # rank 0 prints ONE JSON line on stdout: NCCL writes its version banner to file descriptor 1 when the first
        # communicator comes up, so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    from mri2speech_b200 import _lib, synth
    from mri2speech_b200.pipeline import gather_planned, shard_utterances

    peaks, peak_src = measured_peaks()
    if args.workload == "vocoder":
        v = vocoder_side(device, args.precision, args.steps, peaks)
        if rank == 0:
            print(json.dumps({"metric": "audio_seconds_per_second", "value": v["value"] * world, "unit": "audio-s/s",
                              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": v["ms_per_step"],
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
                              "data": "synthetic", "config": workload_config(args, world), "roofline": v["roofline"]}), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return

    pipe = build_pipeline(device, args.precision, args.max_batch_frames)
    base_host = [c.pin_memory() for c in base_clips()]
    base_dev = [c.to(device) for c in base_host]

    if args.workload == "e2e":
        lengths = synth.synthetic_lengths(64)
        waves = [list(range(64))]
    else:
        lengths = synth.synthetic_lengths(JOB_CLIPS)
        waves = [list(range(w * WAVE_CLIPS, (w + 1) * WAVE_CLIPS)) for w in range(JOB_CLIPS // WAVE_CLIPS)]
    # per wave: this rank's clip ids (LPT on frames) and every rank's total samples (the gather plan: all ranks know the
    # lengths up front, so the gather needs no metadata exchange)
    plans = []
    for ids in waves:
        shards = shard_utterances([lengths[i] for i in ids], world)
        mine = [ids[k] for k in shards[rank]]
        per_rank = [sum(lengths[ids[k]] for k in s) * HOP for s in shards]
        plans.append({"ids": mine, "samples_per_rank": per_rank, "frames": sum(lengths[i] for i in ids),
                      "my_frames": sum(lengths[i] for i in mine)})
    if args.workload == "e2e":
        # configs[2]: the 64 clips as 64 separate pinned host buffers, as a feeder would hold them
        host_clips = [base_host[i % N_BASE][: lengths[i]].clone().pin_memory() for i in plans[0]["ids"]]
        dev_clips = [c.to(device) for c in host_clips]
    gather_bufs = {}

    def run_wave(k, host_inputs, to_host):
        plan = plans[k % len(plans)]
        if args.workload == "e2e":
            clips = host_clips if host_inputs else dev_clips
        else:
            clips = clip_views(base_host if host_inputs else base_dev, lengths, plan["ids"])
        out = pipe.infer(clips, max_batch_frames=args.max_batch_frames,
                         audio_to_host=(to_host and dist is None))
        if dist is not None:
            flat = torch.cat([o["audio"] for o in out]) if out else torch.empty(0, device=device)
            got = gather_planned(flat, plan["samples_per_rank"], dst=0, cache=gather_bufs)
            if to_host and rank == 0:
                key = ("host", k % len(plans))
                if key not in gather_bufs:
                    gather_bufs[key] = [torch.empty(n, dtype=torch.float32).pin_memory() for n in plan["samples_per_rank"]]
                for h, g in zip(gather_bufs[key], got):
                    h.copy_(g[: h.numel()], non_blocking=True)
        return out

    for k in range(args.warmup):
        run_wave(k, False, False)
    torch.cuda.synchronize()

    # ---- timed region (value): inputs resident in HBM ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    _lib.launch_count(reset=True)
    _lib.profile(True)
    total_ms, t0, t1 = timed(lambda k: run_wave(k, False, False), args.steps, dist, device)
    lms, lfl, ltg = _lib.profile_read(with_tags=True)
    _lib.profile(False)
    launches = _lib.launch_count(reset=True)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    frames_total = sum(plans[k % len(plans)]["frames"] for k in range(args.steps))
    my_frames = sum(plans[k % len(plans)]["my_frames"] for k in range(args.steps))
    audio_s_total = frames_total * HOP / SR
    ms_per_step = total_ms / args.steps
    value = audio_s_total / (total_ms * 1e-3)

    # ---- e2e: pinned host clips in, pinned host waveforms out, copies inside the timed region ----
    for k in range(2):
        run_wave(k, True, True)
    e2e_ms, _, _ = timed(lambda k: run_wave(k, True, True), args.steps, dist, device)
    e2e_value = audio_s_total / (e2e_ms * 1e-3)
    h2d = frames_total * 256 * 256 // args.steps
    d2h = frames_total * HOP * 4 // args.steps

    # ---- side measurements (N = 1): the other build on the same workload, the vocoder-only configuration ----
    other_build = vocoder = None
    if world == 1 and not args.no_extras:
        other = "tf32" if args.precision == "fp16" else "fp16"
        del pipe
        torch.cuda.empty_cache()
        pipe2 = build_pipeline(device, other, args.max_batch_frames)
        for _ in range(2):
            pipe2.infer(dev_clips, max_batch_frames=args.max_batch_frames)
        k2 = max(2, min(args.steps, 3))
        ms2, _, _ = timed(lambda k: pipe2.infer(dev_clips, max_batch_frames=args.max_batch_frames), k2)
        other_build = {"precision": other, "ms_per_step": ms2 / k2, "steps": k2, "unit": "audio-s/s",
                       "value": plans[0]["frames"] * HOP / SR / (ms2 / k2 * 1e-3)}
        del pipe2
        torch.cuda.empty_cache()
        vocoder = vocoder_side(device, args.precision, 10, peaks)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the encoder's tcgen05 launches (conv_engine_kernel family) ----
    bd = breakdown(lms, lfl, ltg, args.steps)
    enc = bd.get("encoder_gemm", {"ms_per_step": 0.0, "launch_groups_per_step": 0.0})
    enc_launches = max(enc["launch_groups_per_step"], 1.0)
    my_frames_per_step = my_frames / args.steps
    alg_flop_per_launch = ENC_GEMM_FLOP * my_frames_per_step / enc_launches
    avg_launch_ms = enc["ms_per_step"] / enc_launches
    achieved = alg_flop_per_launch / (avg_launch_ms * 1e-3) / 1e12 if avg_launch_ms > 0 else 0.0
    timed_s = total_ms * 1e-3
    peak_key = "bf16_tflops_sustained" if timed_s >= 2.0 else "bf16_tflops"
    peak = float(peaks.get(peak_key, peaks.get("bf16_tflops")))
    if args.precision == "tf32":
        peak *= 0.5
    kernel_ms = sum(v["ms_per_step"] for v in bd.values())
    traffic = traffic_src = None
    tpath = os.path.join(ROOT, "profiles", "traffic_encoder_engine.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
    roofline = {
        "kernel": f"the encoder's tcgen05 launches: mb_expand_dw_kernel / mb_project_kernel / fused_er_kernel / conv_engine_kernel / "
                  f"conv_engine_pair_kernel (kind::{'f16' if args.precision == 'fp16' else 'tf32'} implicit-GEMM convs of the "
                  f"frame-CNN; the fused kernels also carry the depthwise / squeeze-excite work of their blocks: "
                  f"{100 * enc['ms_per_step'] / max(kernel_ms, 1e-9):.0f} % of the step's kernel time)",
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": f"{peak_src}: dense bf16 cuBLAS, {'sustained' if peak_key.endswith('sustained') else 'burst'} "
                       f"(timed region {timed_s:.1f} s)" + ("; x 0.5 for kind::tf32" if args.precision == "tf32" else ""),
        "algorithmic_flop_per_launch": alg_flop_per_launch, "avg_launch_ms": avg_launch_ms,
        "launches_per_step": enc_launches,
        "path": {"algorithmic_tflop_per_step": PATH_FLOP * my_frames_per_step / 1e12,
                 "achieved_tflops": PATH_FLOP * my_frames_per_step / (ms_per_step * 1e-3) / 1e12,
                 "frac_of_peak": PATH_FLOP * my_frames_per_step / (ms_per_step * 1e-3) / 1e12 / peak,
                 "us_per_frame": ms_per_step * 1e3 / max(my_frames_per_step, 1),
                 "kernel_ms_per_step": kernel_ms, "kernel_share_of_step": kernel_ms / ms_per_step},
        "breakdown_rank0": bd,
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        chain = CpuChain()
        chain.run(chain.clip(4))
        clip = chain.clip(150)
        times = []
        while not times or (sum(times) < 12.0 and len(times) < 5):
            times.append(chain.run(clip))
        rate = 150 * HOP / SR / statistics.median(times)
        cpu = {"value": rate, "unit": "audio-s/s", "cores": chain.threads, "kind": "port",
               "sample": f"BASELINE.json configs[0]: one 150-frame clip end to end (oracle port: restated encoder + nn.LSTM + "
                         f"restated Generator, fp32, B=1), median of {len(times)} passes ({sum(times):.1f} s CPU wall)"}

    line = {
        "metric": "audio_seconds_per_second", "value": value, "unit": "audio-s/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if args.workload == "sharded" else "weak", "vs_baseline": None, "dtype": args.precision,
        "data": "synthetic", "config": workload_config(args, world), "clocks": clocks,
        "frames_per_step": frames_total / args.steps, "us_per_frame_per_gpu": ms_per_step * 1e3 / max(my_frames_per_step, 1),
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps,
                "api": "pipeline.MriToSpeech.infer(pinned uint8 clips, audio_to_host=True)" +
                       (" + gather to rank 0 + D2H on rank 0" if world > 1 else "")},
        "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu, "other_build": other_build, "vocoder": vocoder,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
