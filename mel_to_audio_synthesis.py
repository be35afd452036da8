#!/usr/bin/env python
"""Mel (.npy) -> waveform with the sm_100a HiFi-GAN Generator.  Same flags and output files as the reference CLI
(mel_to_audio_synthesis.py:138-146, :101-135): <name>_from_mel.wav, <name>_input_mel.png (when matplotlib is
installed), <name>_synthesis_stats.json, overall_synthesis_stats.json.  Unlike the reference, which runs one
forward per file, all files go through ONE ragged batch (Generator.forward(..., lengths=...)): every file still
equals its own B=1 result."""
import argparse
import json
import os

import numpy as np
import torch

from models import Generator
from mri2speech_b200 import io_formats


class AttrDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def load_checkpoint(filepath, device):
    assert os.path.isfile(filepath)
    print(f"Loading '{filepath}'...")
    ckpt = torch.load(filepath, map_location=device)
    print("Complete.")
    return ckpt


def collect_mel_files(inp: str, max_files: int):
    if os.path.isfile(inp) and inp.endswith(".npy"):
        return [inp]
    if os.path.isdir(inp):
        files = [os.path.join(inp, f) for f in os.listdir(inp) if f.lower().endswith(".npy")]
        if len(files) > max_files:
            print(f"Found {len(files)} files, processing {max_files} files")
            files = files[:max_files]
        return files
    return []


def synthesize(mel_files, h, generator, device, output_dir):
    """Ragged batch over all files; returns [(basename, stats)]."""
    mels, names = [], []
    for path in mel_files:
        base = os.path.splitext(os.path.basename(path))[0]
        if base.endswith("_mel"):
            base = base[:-4]
        try:
            if not os.path.exists(path):
                raise FileNotFoundError(f"Mel file not found: {path}")
            m = io_formats.fit_mel_bins(io_formats.mel_file_to_tensor(np.load(path)), h.num_mels)
        except Exception as exc:  # the reference reports and skips a bad file (:126-130)
            print(f"Error processing {path}: {exc}")
            continue
        mels.append(m[0])
        names.append((path, base))
    if not mels:
        return []
    lens = [int(m.shape[1]) for m in mels]
    batch = torch.zeros(len(mels), h.num_mels, max(lens), dtype=torch.float32)
    for b, m in enumerate(mels):
        batch[b, :, :lens[b]] = m
    with torch.no_grad():
        audio = generator(batch.to(device), lengths=torch.tensor(lens, dtype=torch.int32))
    hop = audio.shape[-1] // batch.shape[-1]
    results = []
    with io_formats.AsyncWriter() as writer:
        for b, (path, base) in enumerate(names):
            wav = audio[b, 0, : lens[b] * hop]
            writer.submit(wav, lambda a, p=os.path.join(output_dir, f"{base}_from_mel.wav"):
                          io_formats.write_wav_pcm16(p, a, h.sampling_rate))
            lo, hi = float(wav.min()), float(wav.max())
            stats = {"input_file": path, "mel_shape": [1, h.num_mels, lens[b]],
                     "mel_range": [float(mels[b].min()), float(mels[b].max())], "audio_shape": [lens[b] * hop],
                     "audio_range": [lo, hi], "duration_seconds": lens[b] * hop / h.sampling_rate,
                     "sampling_rate": h.sampling_rate}
            with open(os.path.join(output_dir, f"{base}_synthesis_stats.json"), "w") as f:
                json.dump(stats, f, indent=2)
            try:
                import matplotlib
                matplotlib.use("Agg")
                import matplotlib.pyplot as plt
                plt.figure(figsize=(12, 4))
                plt.imshow(mels[b].numpy(), aspect="auto", origin="lower")
                plt.colorbar()
                plt.title(f"Input Mel Spectrogram - {base}")
                plt.tight_layout()
                plt.savefig(os.path.join(output_dir, f"{base}_input_mel.png"), dpi=150)
                plt.close()
            except ImportError:
                pass
            results.append((base, stats))
    return results


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--input", required=True, help="Input .npy mel file or directory with .npy files")
    parser.add_argument("--checkpoint_file", required=True, help="Generator checkpoint file")
    parser.add_argument("--config", default="config_custom.json", help="HiFi-GAN config file")
    parser.add_argument("--output_dir", default="mel_synthesis_result", help="Output directory")
    parser.add_argument("--max_files", default=20, type=int, help="Maximum number of files to process (if directory)")
    args = parser.parse_args(argv)
    with open(args.config) as f:
        h = AttrDict(json.loads(f.read()))
    os.makedirs(args.output_dir, exist_ok=True)
    if not torch.cuda.is_available():
        raise RuntimeError("this build runs on sm_100 CUDA devices only (there is no CPU fallback)")
    device = torch.device("cuda")
    print(f"Using device: {device}")
    mel_files = collect_mel_files(args.input, args.max_files)
    if not mel_files:
        print(f"Invalid input or no .npy files: {args.input}")
        return []
    generator = Generator(h).to(device)
    generator.load_state_dict(load_checkpoint(args.checkpoint_file, device)["generator"])
    generator.eval()  # weight-norm is folded inside libm2s; no removal pass is needed
    results = synthesize(mel_files, h, generator, device, args.output_dir)
    print("\n=== Processing Complete ===")
    print(f"Successfully processed: {len(results)}/{len(mel_files)} files")
    overall = {"total_files": len(mel_files), "successful_syntheses": len(results),
               "model_config": {k: h[k] for k in ("num_mels", "sampling_rate", "n_fft", "hop_size", "win_size")},
               "individual_stats": [s for _, s in results]}
    with open(os.path.join(args.output_dir, "overall_synthesis_stats.json"), "w") as f:
        json.dump(overall, f, indent=2)
    print(f"Results saved to: {args.output_dir}")
    return results


if __name__ == "__main__":
    main()
