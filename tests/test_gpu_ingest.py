"""Loader path on the GPU, through the C ABI: the ingest kernel is bit-exact with the
reference's loader (golden c6_ingest.npz) and with the numpy oracle at the full 30 s geometry; the dataset-file flavour of
K1 matches the float64 oracle to within one float16 ulp."""
import numpy as np
import pytest
import torch

from oracle import ingest_oracle as IO
from oracle import maest_oracle as O

pytestmark = pytest.mark.gpu
NORM_MEAN, NORM_STD = 2.06755686098554, 1.268292820667291


def _bits(t):
    return t.cpu().numpy().view(np.uint16)


def test_ingest_matches_reference_golden(tmp_path, golden):
    from maest_b200 import ingest
    g = golden["c6_ingest"]
    T = int(g["clip_length"]) * 16000 // 256
    cases = [(int(o), None if int(s) == -999 else int(s)) for o, s in g["cases"]]
    fa, fb = str(tmp_path / "a.mmap"), str(tmp_path / "b.mmap")
    g["raw"].tofile(fa)
    g["raw_short"].tofile(fb)
    # load only / load + norm (no roll), all cases + the short file in one batch
    for norm, key in ((False, "load"), (True, "norm")):
        bt = ingest.MelWindowBatcher(batch_size=8, clip_length=int(g["clip_length"]), norm=norm)
        out = bt([fa] * len(cases) + [fb], offsets=[o for o, _ in cases] + [0])
        assert out.shape == (len(cases) + 1, 1, 96, T) and out.dtype == torch.float16
        for i in range(len(cases)):
            assert np.array_equal(_bits(out[i]), g[f"{key}_{i}"].view(np.uint16)), (key, i)
        assert np.array_equal(_bits(out[-1]), g[f"short_{key}"].view(np.uint16))
    # load + norm + roll, one fixed shift per case
    for i, (off, sf) in enumerate(cases):
        if sf is None:
            continue
        bt = ingest.MelWindowBatcher(batch_size=1, clip_length=int(g["clip_length"]), roll=True, roll_shift=sf)
        assert np.array_equal(_bits(bt([fa], offsets=[off])[0]), g[f"roll_{i}"].view(np.uint16)), i


def test_ingest_full_geometry_vs_oracle():
    from maest_b200 import ops
    rng = np.random.RandomState(7)
    B, T = 6, 1875                                   # 30 s windows (discogs/dataset.py:52)
    frames = [1875, 1875, 1000, 1, 1874, 1875]
    shifts = [0, 50, -50, 3, 1874, -1875]
    raw = np.zeros((B, T, 96), dtype=np.float16)
    for b in range(B):
        raw[b, :frames[b]] = (rng.rand(frames[b], 96) * 5).astype(np.float16)
        raw[b, frames[b]:] = np.float16(np.nan)      # rows past frames_read must never be read
    out = ops.mel_ingest(torch.from_numpy(raw).cuda(), torch.tensor(frames, dtype=torch.int32).cuda(),
                         torch.tensor(shifts, dtype=torch.int32).cuda(), NORM_MEAN, NORM_STD)
    for b in range(B):
        exp = IO.ingest(raw[b, :frames[b]], T, 0, NORM_MEAN, NORM_STD, shifts[b])
        assert np.array_equal(_bits(out[b]), exp.view(np.uint16)), b
    # idempotence-style property: no norm, no roll, full windows == plain transpose
    full = (rng.rand(3, T, 96) * 5).astype(np.float16)
    out = ops.mel_ingest(torch.from_numpy(full).cuda())
    assert np.array_equal(_bits(out[:, 0]), np.ascontiguousarray(full.transpose(0, 2, 1)).view(np.uint16))


def test_raw_logmel_file_format(tmp_path):
    from maest_b200 import extract, ops, synth
    x = synth.wave_a(2, 48000)
    got = ops.logmel_raw16(x.cuda()).cpu()
    ref = O.logmel(x, dtype=torch.float64, normalise=False)           # [2, T, 96]
    assert got.shape == ref.shape and got.dtype == torch.float16
    ref16 = ref.to(torch.float16)
    ulp = (got.view(torch.int16).int() - ref16.view(torch.int16).int()).abs()
    assert int(ulp.max()) <= 1 and float((ulp == 0).float().mean()) > 0.995
    # and it is the same spectrogram the model front-end sees, before normalisation
    mel = ops.logmel(x.cuda()).cpu()                                   # [2, 96, T] normalised fp32
    back = (got.float().transpose(1, 2) - NORM_MEAN) / (2 * NORM_STD)
    assert float((back - mel).abs().max()) < 2e-3
    # end-to-end tool: .npy in, raw float16 [frames, 96] .mmap out, centre-trimmed
    wav = str(tmp_path / "a.npy")
    np.save(wav, x[0].numpy())
    dst = str(tmp_path / "out" / "a.mmap")
    shape = extract.main(wav, dst, max_duration=2, framing="torchaudio")
    assert shape == (124, 96)                                           # int(2 * 16000 / 256) = 125 -> 2 * (125 // 2) frames
    disk = np.memmap(dst, dtype="float16", mode="r", shape=shape)
    a, b = IO.trim_bounds(got.shape[1], 2)
    assert np.array_equal(np.asarray(disk).view(np.uint16), got[0, a:b].numpy().view(np.uint16))
    assert extract.main(wav, dst, max_duration=2) is None              # exists, not forced


def test_raw_logmel_essentia_framing(tmp_path):
    """framing="essentia" (the reference's offline extractor, helpers/melspectrogram_extractor.py:15-30): zero-padded centred
    frames, symmetric Hann, ceil(S / 256) frames -- against the float64 restatement of Essentia's published algorithms."""
    from maest_b200 import extract, ops, synth
    for S in (48000, 480000, 4999, 257):
        x = torch.cat([synth.wave_a(1, S), synth.wave_b(S).reshape(1, -1)], 0)
        got = ops.logmel_raw16(x.cuda(), framing="essentia").cpu()
        ref = O.logmel_essentia_framing(x, dtype=torch.float64)
        assert got.shape == ref.shape == (2, (S + 255) // 256, 96)
        ulp = (got.view(torch.int16).int() - ref.to(torch.float16).view(torch.int16).int()).abs()
        assert int(ulp.max()) <= 1 and float((ulp == 0).float().mean()) > 0.99, (S, int(ulp.max()))
    wav = str(tmp_path / "b.npy")
    np.save(wav, synth.wave_a(1, 480000)[0].numpy())
    assert extract.main(wav, str(tmp_path / "b.mmap")) == (1875, 96)   # a 30 s file -> the 1875 frames of the discogs models
    with pytest.raises(ValueError):
        ops.logmel_raw16(x.cuda(), framing="librosa")
