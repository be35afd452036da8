"""CPU, world_size 2 over gloo: the N>1 host logic -- LPT utterance sharding and the final gather of ragged
waveforms (the only collective of the path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lengths, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mri2speech_b200.pipeline import gather_waveforms, shard_utterances
    shards = shard_utterances(lengths, world)
    mine = shards[rank]
    # stand-in "waveforms": deterministic ramps whose content encodes the clip id
    local = [torch.arange(lengths[i] * 4, dtype=torch.float32) + 1000.0 * i for i in mine]
    got = gather_waveforms(local, mine, dst=0)
    if rank == 0:
        ok = sorted(got.keys()) == list(range(len(lengths)))
        for i, ln in enumerate(lengths):
            ok = ok and got[i].numel() == ln * 4 and torch.equal(got[i], torch.arange(ln * 4, dtype=torch.float32) + 1000.0 * i)
        out_q.put(ok)
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_sharding_is_balanced_and_complete():
    from mri2speech_b200.pipeline import shard_utterances
    from mri2speech_b200.synth import synthetic_lengths
    lengths = synthetic_lengths(4096)
    for world in (1, 2, 4, 8):
        shards = shard_utterances(lengths, world)
        assert sorted(i for s in shards for i in s) == list(range(4096))
        loads = [sum(lengths[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(lengths)            # LPT bound
        assert max(loads) / (sum(loads) / world) < 1.001


def test_shard_edge_cases():
    from mri2speech_b200.pipeline import shard_utterances
    assert shard_utterances([], 2) == [[], []]
    assert shard_utterances([5], 4) == [[0], [], [], []]
    s = shard_utterances([3, 3, 3, 3], 2)
    assert sorted(map(len, s)) == [2, 2]


@pytest.mark.timeout(120)
def test_gather_waveforms_world2_gloo():
    lengths = [7, 3, 11, 5, 2]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=90)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert ok


def _worker_planned(rank, world, port, lengths, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mri2speech_b200.pipeline import gather_planned, shard_utterances
    hop = 4
    shards = shard_utterances(lengths, world)
    per_rank = [sum(lengths[i] for i in s) * hop for s in shards]
    cache = {}
    ok = True
    for rep in range(2):                                   # second call reuses the cached buffers
        local = [torch.arange(lengths[i] * hop, dtype=torch.float32) + 1000.0 * i + rep for i in shards[rank]]
        flat = torch.cat(local) if local else torch.empty(0)
        got = gather_planned(flat, per_rank, dst=0, cache=cache)
        if rank == 0:
            for r in range(world):
                off = 0
                for i in shards[r]:
                    n = lengths[i] * hop
                    ok = ok and torch.equal(got[r][off:off + n], torch.arange(n, dtype=torch.float32) + 1000.0 * i + rep)
                    off += n
                ok = ok and off == got[r].numel()
        else:
            assert got is None
    if rank == 0:
        bad = False
        try:
            gather_planned(torch.zeros(3), [1, 1], dst=0)      # local size disagrees with the plan
        except ValueError:
            bad = True
        out_q.put(ok and bad)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_planned_world2_gloo():
    """bench.py's N>1 step: every rank knows the plan, one padded gather, no metadata exchange."""
    lengths = [7, 3, 11, 5, 2, 9]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_planned, args=(r, 2, port, lengths, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=90)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert ok


def test_micro_batch_plan_covers_every_clip_once():
    from mri2speech_b200.pipeline import MriToSpeech
    from mri2speech_b200.synth import synthetic_lengths
    lengths = synthetic_lengths(64)
    plan = MriToSpeech.plan_micro_batches(lengths, 4096)
    assert sorted(i for mb in plan for i in mb) == list(range(64))
    for mb in plan:
        assert len(mb) * max(lengths[i] for i in mb) <= 4096 or len(mb) == 1
    assert MriToSpeech.plan_micro_batches([700], 512) == [[0]]        # a clip longer than the budget still runs
