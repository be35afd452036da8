"""BASELINE.json configs[3] / configs[4]: utterance-sharded rtMRI -> wav throughput with the final gather.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py \
        [--clips 4096] [--precision fp16|tf32] [--sweep] [--alphas 11] [--float-frames]

configs[3] (default): clips (lengths U{150..600}, seed 4321) are dealt to ranks by LPT on frame counts; every rank
runs its shard through MriToSpeech.infer in ragged micro-batches on raw uint8 frames (fused device ingest); the
waveforms are gathered to rank 0 (lengths all_gather + one padded gather).
configs[4] (--sweep): the first `--clips` clips (512 in BASELINE.json) are re-inferred under every (mask, alpha) of the
articulator sweep -- lip / tongue x alpha 0.0..1.0 -- with the mask applied in memory on the device; alpha = 1 (the
identity mask for both presets) is run once.  The final gather covers the last sweep point only (the product's sweep
driver keeps results on the owning rank).
Device timing (CUDA events), max over ranks; rank 0 prints one JSON line."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mri2speech_b200 import masking, synth
from mri2speech_b200.acoustic import build_acoustic_model
from mri2speech_b200.pipeline import MriToSpeech, gather_waveforms, shard_utterances
from mri2speech_b200.vocoder import Generator


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=4096)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "tf32"])
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--alphas", type=int, default=11)
    ap.add_argument("--float-frames", action="store_true", help="float32 frames in [0,1] instead of raw uint8")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h = json.load(open(os.path.join(os.path.dirname(__file__), "..", "config_custom.json")))
    torch.manual_seed(1234)
    pipe = MriToSpeech(build_acoustic_model(precision=args.precision), Generator(h, precision=args.precision),
                       *synth.synthetic_scaler(), device=dev)
    lens = synth.synthetic_lengths(args.clips)
    mine = shard_utterances(lens, world)[rank]
    g = torch.Generator(device=dev).manual_seed(100 + rank)

    def make(ln):
        if not args.float_frames:
            return torch.randint(0, 256, (ln, 256, 256), device=dev, generator=g, dtype=torch.uint8)
        x = torch.rand(ln, 256, 256, device=dev, generator=g)
        mn, mx = x.amin((1, 2), keepdim=True), x.amax((1, 2), keepdim=True)
        return (x - mn) / (mx - mn)

    # sweep points: (mask tensor or None); alpha = 1 is the identity for every preset -> one unmasked run
    points = [None]
    if args.sweep:
        alphas = masking.sweep_alphas(args.alphas)
        points = [torch.from_numpy(masking.preset_mask(mt, a)) for mt in ("lip", "tongue") for a in alphas if a < 1.0]
        points.append(None)

    pipe.reserve(max_batch_frames=8192, max_clip_frames=600)   # workspaces sized once, outside the timed region
    warm = pipe.infer([make(150), make(600)], max_batch_frames=8192)  # warm-up (plans, lazy kernel attributes)
    if world > 1:  # NCCL sets up its point-to-point channels on the first gather: keep that out of the timed region
        gather_waveforms([warm[0]["audio"]], [rank], dst=0)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    group = 32  # clips generated / inferred per round so that frames never exceed a few GB
    waves, ids = [], []
    for i in range(0, len(mine), group):
        idx = mine[i:i + group]
        clips = [make(lens[j]) for j in idx]
        for k, mask in enumerate(points):
            out = pipe.infer(clips, max_batch_frames=8192, mask=mask)
            if k == len(points) - 1:
                waves += [o["audio"] for o in out]
                ids += idx
    gathered = gather_waveforms(waves, ids, dst=0) if world > 1 else dict(zip(ids, waves))
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        assert sorted(gathered.keys()) == list(range(args.clips))
        assert all(gathered[i].numel() == lens[i] * 420 for i in range(0, args.clips, max(1, args.clips // 64)))
        frames = sum(lens) * len(points)
        audio_s = frames * 420 / 11413
        name = (f"configs[4]: articulator sweep, {args.clips} clips x {len(points)} (mask, alpha) points "
                f"(2 presets x {args.alphas} alphas, alpha=1 deduplicated)" if args.sweep else f"configs[3]: {args.clips} clips")
        print(json.dumps({"workload": f"{name} ({frames} frames, {audio_s:.0f} s audio) sharded over {world} GPU(s), LPT, "
                                      "final gather", "n_gpus": world, "precision": args.precision,
                          "frames_dtype": "float32" if args.float_frames else "uint8", "ms": float(ms.item()),
                          "audio_s_per_s": audio_s / (float(ms.item()) * 1e-3),
                          "us_per_frame_per_gpu": float(ms.item()) * 1e3 * world / frames,
                          "note": "timed region includes synthetic frame generation on the device"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
