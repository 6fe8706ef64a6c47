"""Loader -> device ingest: the host side of `maest_mel_ingest_fwd`.

Mirrors `DiscogsDataset.load_melspectrogram` (discogs/dataset.py:88-139) + the data module's `norm_func` / `roll_func`
(discogs/datamodule.py:111-137), split the B200 way: the host only decides WHICH bytes of each raw float16 `[frames, 96]`
`.mmap` file form the window (same `random.randint` / `np.random.random_integers` draws as the reference, in the same
order) and copies those bytes, untouched and time-major, into one pinned batch buffer; zero-padding, the centring roll,
the transpose to `[1, 96, T]`, the float16 normalisation and the time roll happen in one kernel on the GPU.
"""
from __future__ import annotations

import os
import random
from typing import Optional, Sequence

import numpy as np
import torch

from . import ops

N_BANDS = 96
NORM_MEAN, NORM_STD = 2.06755686098554, 1.268292820667291     # discogs/datamodule.py:49-53


def window_of(frames_num: int, melspectrogram_size: int, offset: Optional[int] = None):
    """(offset, frames_to_read) of one clip, as discogs/dataset.py:93-102 computes them (random offset when None)."""
    if type(offset) is not int:
        max_frame = frames_num - melspectrogram_size
        offset = random.randint(0, max(max_frame, 0))
    skip_frames = max(offset + melspectrogram_size - frames_num, 0)
    return offset, melspectrogram_size - skip_frames


class MelWindowBatcher:
    """Reads windows of raw float16 mel files into a pinned `[B, T, 96]` buffer and runs the ingest kernel.

    `clip_length` / `sample_rate` / `hop_size` give `melspectrogram_size` exactly as discogs/dataset.py:52 does."""

    def __init__(self, batch_size: int, clip_length: int, sample_rate: int = 16000, hop_size: int = 256, device="cuda",
                 norm: bool = True, norm_mean: float = NORM_MEAN, norm_std: float = NORM_STD,
                 roll: bool = False, roll_shift: Optional[int] = None, roll_shift_range: int = 50):
        self.T = clip_length * sample_rate // hop_size
        self.B = batch_size
        self.device = torch.device(device)
        self.norm = (norm_mean, norm_std) if norm else (None, None)
        self.roll, self.shift, self.shift_range = roll, roll_shift, roll_shift_range
        pin = self.device.type == "cuda"
        self.host = torch.zeros((batch_size, self.T, N_BANDS), dtype=torch.float16, pin_memory=pin)
        self.host_n = torch.zeros(batch_size, dtype=torch.int32, pin_memory=pin)
        self.host_s = torch.zeros(batch_size, dtype=torch.int32, pin_memory=pin)
        self._h2d_done = None      # CUDA event recorded after the last batch's H2D copies (the pinned buffers are reused)

    def stage(self, files: Sequence[str], offsets: Optional[Sequence[Optional[int]]] = None):
        """Host half: copy each file's window bytes into the pinned buffer; returns (n_clips, frames_read, shifts)."""
        assert len(files) <= self.B
        if self._h2d_done is not None:
            # the previous batch's asynchronous H2D copies may still be queued behind earlier kernels: do not overwrite the
            # pinned staging buffers until the DMA has actually read them
            self._h2d_done.synchronize()
            self._h2d_done = None
        buf = self.host.numpy()
        for i, f in enumerate(files):
            if str(f).endswith(".npy"):
                # discogs/dataset.py:72-87: whole-array files; no offset draw, the first T frames (or all of them, padded)
                arr = np.load(f).astype("float16")
                n = min(arr.shape[0], self.T)
                buf[i, :n] = arr[:n]
                self.host_n[i] = n
                continue
            frames_num = os.stat(f).st_size // (2 * N_BANDS)
            off, n = window_of(frames_num, self.T, None if offsets is None else offsets[i])
            fp = np.memmap(f, dtype="float16", mode="r", shape=(n, N_BANDS), offset=off * N_BANDS * 2)
            buf[i, :n] = fp
            del fp
            self.host_n[i] = n
        if self.roll:   # one draw per clip, after the window draw, as the wrapped datasets do (datamodule.py:116-122)
            for i in range(len(files)):
                sf = self.shift
                if sf is None:
                    sf = int(np.random.randint(-self.shift_range, self.shift_range + 1))   # == random_integers(-r, r)
                self.host_s[i] = sf
        return len(files)

    def to_device(self, n_clips: int) -> torch.Tensor:
        """Device half: H2D of the raw windows + ONE kernel -> `[n, 1, 96, T]` float16 (what `Module.training_step` takes)."""
        raw = self.host[:n_clips].to(self.device, non_blocking=True)
        nread = self.host_n[:n_clips].to(self.device, non_blocking=True)
        shift = self.host_s[:n_clips].to(self.device, non_blocking=True) if self.roll else None
        if self.device.type == "cuda":
            self._h2d_done = torch.cuda.Event()
            self._h2d_done.record(torch.cuda.current_stream(self.device))
        return ops.mel_ingest(raw, nread, shift, self.norm[0], self.norm[1])

    def __call__(self, files, offsets=None) -> torch.Tensor:
        return self.to_device(self.stage(files, offsets))
