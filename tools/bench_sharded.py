"""BASELINE.json configs[3]: utterance-sharded rtMRI -> wav throughput with the final gather.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sharded.py [n_clips=4096]

Clips (lengths U{150..600}, seed 4321) are dealt to ranks by LPT on frame counts; every rank runs its shard through
MriToSpeech.infer in ragged micro-batches; the waveforms are gathered to rank 0 (lengths all_gather + one padded
gather).  Device timing (CUDA events), max over ranks; rank 0 prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mri2speech_b200 import synth
from mri2speech_b200.acoustic import build_acoustic_model
from mri2speech_b200.pipeline import MriToSpeech, gather_waveforms, shard_utterances
from mri2speech_b200.vocoder import Generator


def main():
    n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h = json.load(open(os.path.join(os.path.dirname(__file__), "..", "config_custom.json")))
    torch.manual_seed(1234)
    pipe = MriToSpeech(build_acoustic_model(), Generator(h), *synth.synthetic_scaler(), device=dev)
    lens = synth.synthetic_lengths(n_clips)
    mine = shard_utterances(lens, world)[rank]
    g = torch.Generator(device=dev).manual_seed(100 + rank)

    def make(ln):
        x = torch.rand(ln, 256, 256, device=dev, generator=g)
        mn, mx = x.amin((1, 2), keepdim=True), x.amax((1, 2), keepdim=True)
        return (x - mn) / (mx - mn)

    pipe.infer([make(150)])  # warm-up (plans, workspaces)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    waves, ids = [], []
    group = 32  # clips generated / inferred per round so that frames never exceed a few GB
    for i in range(0, len(mine), group):
        idx = mine[i:i + group]
        out = pipe.infer([make(lens[j]) for j in idx], max_batch_frames=8192)
        waves += [o["audio"] for o in out]
        ids += idx
    gathered = gather_waveforms(waves, ids, dst=0) if world > 1 else dict(zip(ids, waves))
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        assert sorted(gathered.keys()) == list(range(n_clips))
        assert all(gathered[i].numel() == lens[i] * 420 for i in range(0, n_clips, max(1, n_clips // 64)))
        frames = sum(lens)
        audio_s = frames * 420 / 11413
        print(json.dumps({"workload": f"configs[3]: {n_clips} clips ({frames} frames, {audio_s:.0f} s audio) sharded over "
                                      f"{world} GPU(s), LPT, final gather", "n_gpus": world, "ms": float(ms.item()),
                          "audio_s_per_s": audio_s / (float(ms.item()) * 1e-3),
                          "note": "timed region includes synthetic frame generation on the device"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
