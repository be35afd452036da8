#!/usr/bin/env python
"""Directory of mel .npy files -> int16 wavs (reference inference_e2e.py:33-59, flags :65-69, config.json next to
the checkpoint :71-77), on the sm_100a Generator.  The reference calls Generator.remove_weight_norm(), which raises
on its un-normed conv_pre (models.py:139, SURVEY.md 8f-4 "a fixed inference_e2e.py"); weight-norm is folded inside
libm2s, so no removal pass exists here."""
import argparse
import json
import os

import numpy as np
import torch

from env import AttrDict
from models import Generator
from mri2speech_b200 import io_formats


def load_checkpoint(filepath, device):
    assert os.path.isfile(filepath)
    print("Loading '{}'".format(filepath))
    ckpt = torch.load(filepath, map_location=device)
    print("Complete.")
    return ckpt


def inference(a, h, device):
    generator = Generator(h).to(device)
    generator.load_state_dict(load_checkpoint(a.checkpoint_file, device)["generator"])
    generator.eval()
    filelist = sorted(os.listdir(a.input_mels_dir))
    os.makedirs(a.output_dir, exist_ok=True)
    outputs = []
    with torch.no_grad(), io_formats.AsyncWriter() as writer:
        for name in filelist:
            x = torch.FloatTensor(np.load(os.path.join(a.input_mels_dir, name))).to(device)
            audio = generator(x).squeeze()
            out = os.path.join(a.output_dir, os.path.splitext(name)[0] + "_generated_e2e.wav")
            writer.submit(audio, lambda arr, p=out: io_formats.write_wav_int16(p, arr, h.sampling_rate))
            outputs.append(out)
            print(out)
    return outputs


def main(argv=None):
    print("Initializing Inference Process..")
    parser = argparse.ArgumentParser()
    parser.add_argument("--input_mels_dir", default="test_mel_files")
    parser.add_argument("--output_dir", default="generated_files_from_mel")
    parser.add_argument("--checkpoint_file", required=True)
    a = parser.parse_args(argv)
    with open(os.path.join(os.path.split(a.checkpoint_file)[0], "config.json")) as f:
        h = AttrDict(json.loads(f.read()))
    torch.manual_seed(h.seed)
    if not torch.cuda.is_available():
        raise RuntimeError("this build runs on sm_100 CUDA devices only (there is no CPU fallback)")
    torch.cuda.manual_seed(h.seed)
    return inference(a, h, torch.device("cuda"))


if __name__ == "__main__":
    main()
