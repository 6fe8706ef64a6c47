"""Offline mel extraction at scale: K1 as a batch tool that writes the reference's dataset
file format -- raw float16 `[frames, 96]` `.mmap`, un-normalised log10(1 + 1e4 mel), centre-trimmed to `max_duration`
(helpers/melspectrogram_extractor.py:15-48, datasets/mtt/preprocess.py:44-64).

    python -m maest_b200.extract audio.wav melbands.mmap [--force] [--max-duration 300] [--framing essentia|torchaudio]

The reference runs Essentia (not installed here, not vendored by the reference).  `--framing essentia` (default) reproduces its
framing as published -- centred frames with zero padding, symmetric Hann, ceil(S / 256) frames, so a 30 s file gives the 1875
frames the discogs models were trained on; `--framing torchaudio` runs the model's own front-end (reflect padding, periodic
Hann, 1 + S // 256 frames).  Audio decoding is limited to what the image has: 16 kHz mono PCM `.wav` (stdlib `wave`) or `.npy`
float arrays.
"""
from __future__ import annotations

import argparse
import os
import wave
from pathlib import Path

import numpy as np
import torch

from . import ops

SR, HOP_SIZE, N_MELS = 16000, 256, 96


def load_audio(audio_file: str) -> np.ndarray:
    if audio_file.endswith(".npy"):
        return np.asarray(np.load(audio_file), dtype=np.float32).reshape(-1)
    with wave.open(audio_file, "rb") as w:
        if w.getframerate() != SR:
            raise RuntimeError(f"{audio_file}: {w.getframerate()} Hz; this tool expects {SR} Hz input (no resampler in the image)")
        n, ch, width = w.getnframes(), w.getnchannels(), w.getsampwidth()
        raw = w.readframes(n)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    else:
        raise RuntimeError(f"{audio_file}: unsupported sample width {width}")
    return x.reshape(-1, ch).mean(1) if ch > 1 else x


def trim_bounds(n_frames: int, max_duration: float):
    """Centre trim of helpers/melspectrogram_extractor.py:37-42: returns (first, last) frame kept."""
    max_timestamps = int(max_duration * SR / HOP_SIZE)
    if n_frames > max_timestamps:
        mid = n_frames // 2
        return mid - max_timestamps // 2, mid + max_timestamps // 2
    return 0, n_frames


def melspectrogram_extractor(waveform: torch.Tensor, framing: str = "essentia") -> torch.Tensor:
    """[S] or [B, S] waveform on the GPU -> float16 [T, 96] / [B, T, 96] (time-major, un-normalised)."""
    return ops.logmel_raw16(waveform, framing=framing)


def main(audio_file, melbands_file, force=False, max_duration=300, device="cuda", framing="essentia"):
    if os.path.exists(melbands_file) and not force:
        return None
    x = torch.from_numpy(load_audio(audio_file)).to(device)
    mel = melspectrogram_extractor(x, framing)
    a, b = trim_bounds(mel.shape[0], max_duration)
    mel = mel[a:b].cpu().numpy()
    Path(melbands_file).parent.mkdir(parents=True, exist_ok=True)
    fp = np.memmap(melbands_file, dtype="float16", mode="w+", shape=mel.shape)
    fp[:] = mel[:]
    del fp
    return mel.shape


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Computes the mel spectrogram of a given audio file (B200 K1 kernel).")
    ap.add_argument("audio_file")
    ap.add_argument("melbands_file", type=str)
    ap.add_argument("--force", "-f", action="store_true")
    ap.add_argument("--max-duration", type=float, default=300)
    ap.add_argument("--framing", default="essentia", choices=["essentia", "torchaudio"])
    main(**vars(ap.parse_args()))
