"""Golden fixture for the loader path: runs the UNMODIFIED reference's
`DiscogsDataset.load_melspectrogram` (discogs/dataset.py) and `DiscogsDataModule.get_norm_func / get_roll_func`
(discogs/datamodule.py) on a small seeded raw float16 file.  Dev container only:  python tests/golden/make_golden_ingest.py
Writes tests/golden/c6_ingest.npz (raw file content + the reference outputs per case)."""
from __future__ import annotations

import importlib
import os
import pickle
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

CLIP_LENGTH = 2              # seconds -> melspectrogram_size = 2 * 16000 // 256 = 125 frames
FRAMES = 300
# (offset, roll shift): inside the file / running past the end (zero-pad + centring roll) / whole-file-shorter cases
CASES = [(0, None), (37, 11), (175, -50), (200, 3), (290, 124), (299, -7)]
SHORT_FRAMES = 80            # a file shorter than one window, read at offset 0
NORM_MEAN, NORM_STD = 2.06755686098554, 1.268292820667291


def main():
    ref_loader.install_stubs(with_lightning=True)
    sys.modules["lightning.pytorch"].LightningDataModule = object
    sys.modules["pytorch_lightning"].LightningDataModule = object
    sys.path.insert(0, ref_loader.REF_ROOT)
    dataset = importlib.import_module("discogs.dataset")
    try:
        datamodule = importlib.import_module("discogs.datamodule")
    except Exception as e:   # noqa: BLE001
        raise SystemExit(f"cannot import the reference datamodule: {e!r}")

    rng = np.random.RandomState(20260)
    raw = (rng.rand(FRAMES, 96) * 4.5).astype("float16")          # log10(1 + 1e4 mel) lives in [0, ~5]
    raw_short = (rng.rand(SHORT_FRAMES, 96) * 4.5).astype("float16")
    out = dict(raw=raw, raw_short=raw_short, clip_length=CLIP_LENGTH, cases=np.array([[o, -999 if s is None else s] for o, s in CASES]))
    with tempfile.TemporaryDirectory() as d:
        gt = os.path.join(d, "gt.pk")
        with open(gt, "wb") as f:
            pickle.dump({"a.mmap": np.zeros(400, dtype="float16"), "b.mmap": np.zeros(400, dtype="float16")}, f)
        raw.tofile(os.path.join(d, "a.mmap"))
        raw_short.tofile(os.path.join(d, "b.mmap"))
        ds = dataset.DiscogsDataset(groundtruth_file=gt, base_dir=d, sample_rate=16000, clip_length=CLIP_LENGTH, hop_size=256, n_bands=96)
        dm = types.SimpleNamespace()
        norm_func = datamodule.DiscogsDataModule.get_norm_func(dm, NORM_MEAN, NORM_STD)
        import pathlib
        for i, (off, sf) in enumerate(CASES):
            x = ds.load_melspectrogram(pathlib.Path(d, "a.mmap"), offset=off)
            out[f"load_{i}"] = x
            b = norm_func((x, "a.mmap", None))
            xn = b[0]
            out[f"norm_{i}"] = np.asarray(xn)
            if sf is not None:
                roll_func = datamodule.DiscogsDataModule.get_roll_func(dm, -1, sf, 50)
                out[f"roll_{i}"] = roll_func((xn, "a.mmap", None))[0].numpy()
        x = ds.load_melspectrogram(pathlib.Path(d, "b.mmap"), offset=0)
        out["short_load"] = x
        out["short_norm"] = np.asarray(norm_func((x, "b.mmap", None))[0])
    for k, v in out.items():
        if isinstance(v, np.ndarray) and v.dtype != np.int64:
            assert v.dtype == np.float16, (k, v.dtype)
    np.savez_compressed(os.path.join(HERE, "c6_ingest.npz"), **out)
    print("wrote c6_ingest.npz:", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
