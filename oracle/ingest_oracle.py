"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's loader path for SURVEY.md section 8(f) rows 1-2.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Pinned against outputs of the UNMODIFIED reference run in the dev container (tests/golden/make_golden.py -> c6.npz):
`DiscogsDataset.load_melspectrogram` (discogs/dataset.py:68-139) and the `norm_func` / `roll_func` closures of
`DiscogsDataModule` (discogs/datamodule.py:111-137)."""
from __future__ import annotations

import numpy as np


def load_window(raw: np.ndarray, melspectrogram_size: int, offset: int) -> np.ndarray:
    """raw: the whole file as float16 [frames, 96].  Returns [1, 96, T] float16 (discogs/dataset.py:99-139)."""
    frames_num = raw.shape[0]
    skip_frames = max(offset + melspectrogram_size - frames_num, 0)          # :101
    frames_to_read = melspectrogram_size - skip_frames                      # :102
    mel = np.array(raw[offset:offset + frames_to_read], dtype="float16")    # :105-120
    if frames_to_read < melspectrogram_size:                                # :122-133
        padding_size = melspectrogram_size - frames_to_read
        mel = np.vstack([mel, np.zeros([padding_size, raw.shape[1]], dtype="float16")])
        mel = np.roll(mel, padding_size // 2, axis=0)
    return np.expand_dims(mel.T, 0)                                         # :137-138


def load_npy(arr: np.ndarray, melspectrogram_size: int) -> np.ndarray:
    """The `.npy` branch of load_melspectrogram (discogs/dataset.py:72-87): no window offset; short arrays are zero-padded
    and the padding centred, long ones truncated to the first `melspectrogram_size` frames."""
    mel = arr.astype("float16")
    if mel.shape[0] < melspectrogram_size:
        padding_size = melspectrogram_size - mel.shape[0]
        mel = np.vstack([mel, np.zeros([padding_size, mel.shape[1]], dtype="float16")])
        mel = np.roll(mel, padding_size // 2, axis=0)
    else:
        mel = mel[:melspectrogram_size, :]
    return np.expand_dims(mel.T, 0)


def norm(x: np.ndarray, norm_mean: float, norm_std: float) -> np.ndarray:
    """discogs/datamodule.py:130-134 on a float16 array: numpy keeps float16 (the Python floats are weak scalars)."""
    return (x - norm_mean) / (norm_std * 2)


def roll(x: np.ndarray, shift: int, axis: int = -1) -> np.ndarray:
    """discogs/datamodule.py:116-122: torch.roll == np.roll."""
    return np.roll(x, shift, axis)


def ingest(raw: np.ndarray, melspectrogram_size: int, offset: int, norm_mean=None, norm_std=None, shift=None) -> np.ndarray:
    x = load_window(raw, melspectrogram_size, offset)
    if norm_mean is not None:
        x = norm(x, norm_mean, norm_std)
    if shift is not None:
        x = roll(x, shift, -1)
    return x


def trim_bounds(n_frames: int, max_duration: float, sr: int = 16000, hop: int = 256):
    """helpers/melspectrogram_extractor.py:35-42."""
    max_timestamps = int(max_duration * sr / hop)
    if n_frames > max_timestamps:
        mid = n_frames // 2
        return mid - max_timestamps // 2, mid + max_timestamps // 2
    return 0, n_frames
