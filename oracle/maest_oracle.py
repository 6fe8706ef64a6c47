"""CPU oracle for the MAEST hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may
import this file, and only as the checker / reported CPU baseline.  `maest_b200/` never imports it.

It is an independent restatement (plain torch CPU tensor ops, any float dtype; float64 for pinning)
of the reference's algorithm for: waveform -> log-mel -> patch tokens (+pos-embed, +patchout)
-> 12 transformer blocks -> pooling/head, following
  * models/helpers/melspectrogram.py:16-60          (constants, log compression, z-norm)
  * torchaudio 2.11 (third-party, unpinned in pyproject.toml:17-29; not under /root/reference):
      functional.spectrogram (center=True reflect pad, periodic Hann, onesided, power 2),
      functional.melscale_fbanks(mel_scale="slaney", norm="slaney"), transforms.MelScale
  * models/maest.py:183-208 (Mlp) :214-256 (PatchEmbed) :346-420 (Attention, Block)
    :634-829 (forward_features) :831-939 (forward, predict_labels)
  * models/module.py:73-102 + helpers/mixup.py:5-12  (training step: mixup, BCE-with-logits)

Pinning: the reference has NO golden vectors or known-answer tests (tests/test_maest.py holds shape
and exception checks only).  The oracle is pinned against outputs of the reference itself, run in
the dev container by `tests/golden/make_golden.py`, committed as `tests/golden/*.npz`
(`tests/test_oracle_golden.py`), and — when /root/reference is present — live
(`tests/test_oracle_live_reference.py`).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch

# models/helpers/melspectrogram.py:16-24
SR = 16000
N_FFT = 512
HOP = 256
N_MELS = 96
NORM_MEAN = 2.06755686098554
NORM_STD = 1.268292820667291

PATCH = 16
STRIDE = 10
EMBED = 768
HEADS = 12
HEAD_DIM = 64


# ----------------------------------------------------------------------------------------------
# log-mel front-end
# ----------------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f: float) -> float:
    # torchaudio functional._hz_to_mel(mel_scale="slaney")
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    if f >= min_log_hz:
        return min_log_mel + math.log(f / min_log_hz) / logstep
    return f / f_sp


def _mel_to_hz_slaney(m: torch.Tensor) -> torch.Tensor:
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    lin = f_sp * m
    log = min_log_hz * torch.exp(logstep * (m - min_log_mel))
    return torch.where(m >= min_log_mel, log, lin)


def mel_filterbank(dtype=torch.float64) -> torch.Tensor:
    """fb[257, 96]: triangular Slaney filters with Slaney area normalisation, f in [0, 8000] Hz.

    torchaudio melscale_fbanks(n_freqs=257, f_min=0, f_max=sr/2, n_mels=96, sample_rate=16000,
    norm="slaney", mel_scale="slaney") as instantiated by models/helpers/melspectrogram.py:36-42.
    """
    n_freqs = N_FFT // 2 + 1
    all_freqs = torch.linspace(0, SR // 2, n_freqs, dtype=torch.float64)
    m_min = _hz_to_mel_slaney(0.0)
    m_max = _hz_to_mel_slaney(SR / 2.0)
    m_pts = torch.linspace(m_min, m_max, N_MELS + 2, dtype=torch.float64)
    f_pts = _mel_to_hz_slaney(m_pts)
    f_diff = f_pts[1:] - f_pts[:-1]                      # [97]
    slopes = f_pts[None, :] - all_freqs[:, None]          # [257, 98]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.clamp(torch.minimum(down, up), min=0.0)
    enorm = 2.0 / (f_pts[2: N_MELS + 2] - f_pts[:N_MELS])
    fb = fb * enorm[None, :]
    return fb.to(dtype)


def hann_periodic(dtype=torch.float64) -> torch.Tensor:
    n = torch.arange(N_FFT, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2 * math.pi * n / N_FFT)).to(dtype)


def n_frames(samples: int) -> int:
    return 1 + samples // HOP


def logmel(wave: torch.Tensor, dtype=torch.float32, normalise: bool = True) -> torch.Tensor:
    """[..., S] waveform -> [..., 96, T] normalised log-mel, T = 1 + S // 256.

    models/helpers/melspectrogram.py:47-60.  Frame t = xp[256 t : 256 t + 512] of the waveform
    reflect-padded by 256 samples (edge sample not repeated), times the periodic Hann window;
    |rfft|^2; mel = power . fb; log10(1 + 1e4 mel); (x - mean) / (2 std).
    """
    lead = wave.shape[:-1]
    x = wave.reshape(-1, wave.shape[-1]).to(dtype)
    S = x.shape[-1]
    pad = N_FFT // 2
    left = x[:, 1: pad + 1].flip(-1)
    right = x[:, S - pad - 1: S - 1].flip(-1)
    xp = torch.cat([left, x, right], dim=-1)
    T = n_frames(S)
    frames = xp.unfold(-1, N_FFT, HOP)[:, :T, :]          # [B, T, 512]
    frames = frames * hann_periodic(dtype)
    spec = torch.fft.rfft(frames, dim=-1)
    power = spec.real ** 2 + spec.imag ** 2               # [B, T, 257]
    mel = power @ mel_filterbank(dtype)                   # [B, T, 96]
    out = torch.log10(1.0 + mel * 10000.0)
    if not normalise:      # dataset-file flavour (helpers/melspectrogram_extractor.py): time-major, un-normalised
        return out.reshape(*lead, T, N_MELS)
    out = (out - NORM_MEAN) / (NORM_STD * 2)
    return out.transpose(1, 2).reshape(*lead, N_MELS, T)


def hann_symmetric(dtype=torch.float64) -> torch.Tensor:
    n = torch.arange(N_FFT, dtype=torch.float64)
    return (0.5 - 0.5 * torch.cos(2.0 * math.pi * n / (N_FFT - 1))).to(dtype)


def logmel_essentia_framing(wave: torch.Tensor, dtype=torch.float64) -> torch.Tensor:
    """[..., S] waveform -> [..., T, 96] UN-normalised log10(1 + 1e4 mel), T = ceil(S / 256): the dataset-file flavour with the
    framing of the reference's offline extractor (helpers/melspectrogram_extractor.py:15-30 -> essentia.pytools.extractors.
    melspectrogram).  PARITY UNPINNED: Essentia is a third-party dependency (unpinned in pyproject.toml:17-29), neither installed
    here nor vendored under /root/reference, and the reference holds no vectors produced by it; this restates its published
    algorithms: FrameCutter(frameSize=512, hopSize=256, startFromZero=False): frame t covers samples [256 t - 256, 256 t + 256)
    with zeros outside the signal, frames while the centre 256 t lies inside it; Windowing(type='hann', normalized=False):
    symmetric Hann (zero-phase rotation does not change magnitudes); Spectrum -> |rfft|; MelBands(96 bands, 0-8000 Hz, slaneyMel
    warping, 'unit_tri' normalisation, type='power'): the same triangles as torchaudio's slaney filterbank applied to |X|^2;
    'shift_scale_log': log10(1 + 1e4 x).  Interior frames differ from `logmel(normalise=False)` only by the window
    (periodic vs symmetric Hann); the reference's authors quote 1e-3 relative / 1e-3 absolute between the two
    (models/helpers/melspectrogram.py:8-10) and tests/test_oracle_golden.py::test_essentia_framing_gap measures it here."""
    lead = wave.shape[:-1]
    x = wave.reshape(-1, wave.shape[-1]).to(dtype)
    S = x.shape[-1]
    T = (S + HOP - 1) // HOP
    pad = N_FFT // 2
    need = (T - 1) * HOP + N_FFT
    xp = torch.cat([x.new_zeros(x.shape[0], pad), x, x.new_zeros(x.shape[0], max(0, need - pad - S))], dim=-1)
    frames = xp.unfold(-1, N_FFT, HOP)[:, :T, :] * hann_symmetric(dtype)
    spec = torch.fft.rfft(frames, dim=-1)
    power = spec.real ** 2 + spec.imag ** 2
    mel = power @ mel_filterbank(dtype)
    return torch.log10(1.0 + mel * 10000.0).reshape(*lead, T, N_MELS)


# ----------------------------------------------------------------------------------------------
# tokens
# ----------------------------------------------------------------------------------------------
def patch_grid(n_mel_rows: int, t_frames: int) -> Tuple[int, int]:
    return (n_mel_rows - PATCH) // STRIDE + 1, (t_frames - PATCH) // STRIDE + 1


def patch_tokens(mel: torch.Tensor, sd: Dict[str, torch.Tensor], t_offset: int = 0,
                 keep_t: Optional[Sequence[int]] = None, keep_f: Optional[Sequence[int]] = None,
                 keep_seq: Optional[Sequence[int]] = None) -> torch.Tensor:
    """mel [B,96,T] -> token sequence [B, 2 + P, 768] entering blocks[0].

    models/maest.py:243-256 (Conv2d k=16 s=10 as a patch GEMM, K index = kh*16 + kw),
    :645-675 (+time pos-embed columns [t_offset, t_offset+T'), +freq pos-embed),
    :678-766 (structured patchout = keep sorted column / row subsets, same for every clip),
    :769 (frequency-major flatten p = f*T'_kept + t), :773-778 (unstructured patchout),
    :785-796 (prepend cls+pos0, dist+pos1).
    """
    dt = mel.dtype
    B, Fm, T = mel.shape
    Fp, Tp = patch_grid(Fm, T)
    time_pe = sd["time_new_pos_embed"].to(dt)             # [1,768,1,W]
    if Tp > time_pe.shape[-1]:
        raise Exception("patches wider than the time encodings")  # models/maest.py:664-668
    pt = mel.unfold(1, PATCH, STRIDE).unfold(2, PATCH, STRIDE)    # [B,F',T',16(kh),16(kw)]
    a = pt.reshape(B, Fp, Tp, PATCH * PATCH)
    w = sd["patch_embed.proj.weight"].to(dt).reshape(EMBED, PATCH * PATCH)
    x = a @ w.t() + sd["patch_embed.proj.bias"].to(dt)            # [B,F',T',768]
    x = x + time_pe[0, :, 0, t_offset: t_offset + Tp].t()[None, None, :, :]
    x = x + sd["freq_new_pos_embed"].to(dt)[0, :, :, 0].t()[None, :, None, :]
    if keep_t is not None:
        x = x[:, :, list(keep_t), :]
    if keep_f is not None:
        x = x[:, list(keep_f), :, :]
    x = x.reshape(B, -1, EMBED)
    if keep_seq is not None:
        x = x[:, list(keep_seq), :]
    npe = sd["new_pos_embed"].to(dt)
    cls = (sd["cls_token"].to(dt) + npe[:, :1, :]).expand(B, -1, -1)
    dist = (sd["dist_token"].to(dt) + npe[:, 1:, :]).expand(B, -1, -1)
    return torch.cat([cls, dist, x], dim=1)


# ----------------------------------------------------------------------------------------------
# encoder
# ----------------------------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float) -> torch.Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def attention(h: torch.Tensor, sd, p: str) -> torch.Tensor:
    """models/maest.py:358-378 — qkv rows are [3][12 heads][64]; scale 64^-0.5; softmax over keys."""
    dt = h.dtype
    B, N, C = h.shape
    qkv = h @ sd[p + "attn.qkv.weight"].to(dt).t() + sd[p + "attn.qkv.bias"].to(dt)
    qkv = qkv.reshape(B, N, 3, HEADS, HEAD_DIM).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    s = (q @ k.transpose(-2, -1)) * (HEAD_DIM ** -0.5)
    s = s - s.amax(-1, keepdim=True)
    e = torch.exp(s)
    a = e / e.sum(-1, keepdim=True)
    o = (a @ v).transpose(1, 2).reshape(B, N, C)
    return o @ sd[p + "attn.proj.weight"].to(dt).t() + sd[p + "attn.proj.bias"].to(dt)


def mlp(h: torch.Tensor, sd, p: str) -> torch.Tensor:
    dt = h.dtype
    u = h @ sd[p + "mlp.fc1.weight"].to(dt).t() + sd[p + "mlp.fc1.bias"].to(dt)
    u = gelu_erf(u)
    return u @ sd[p + "mlp.fc2.weight"].to(dt).t() + sd[p + "mlp.fc2.bias"].to(dt)


def block(x: torch.Tensor, sd, i: int, return_self_attention: bool = False) -> torch.Tensor:
    """models/maest.py:414-420 — pre-LN block, LN eps 1e-6 (:499), DropPath = identity."""
    p = f"blocks.{i}."
    dt = x.dtype
    a = attention(layer_norm(x, sd[p + "norm1.weight"].to(dt), sd[p + "norm1.bias"].to(dt), 1e-6), sd, p)
    if return_self_attention:
        return a
    x = x + a
    x = x + mlp(layer_norm(x, sd[p + "norm2.weight"].to(dt), sd[p + "norm2.bias"].to(dt), 1e-6), sd, p)
    return x


def n_blocks(sd) -> int:
    return 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))


def encode(x: torch.Tensor, sd, transformer_block: int = -1, return_self_attention: bool = False):
    """models/maest.py:804-829.  -1: all blocks, final LN(1e-6), rows 0 and 1.
    k>=0: blocks 0..k and cat(x[:,0], x[:,1], mean(x[:,2:],1)) WITHOUT the final LN."""
    depth = n_blocks(sd)
    dt = x.dtype
    if transformer_block == -1:
        for i in range(depth):
            x = block(x, sd, i)
        x = layer_norm(x, sd["norm.weight"].to(dt), sd["norm.bias"].to(dt), 1e-6)
        return x[:, 0], x[:, 1]
    for i in range(depth):
        if i == transformer_block:
            x = block(x, sd, i, return_self_attention=return_self_attention)
            break
        x = block(x, sd, i)
    return torch.cat([x[:, 0, :], x[:, 1, :], x[:, 2:, :].mean(dim=1)], dim=1)


def head(cls: torch.Tensor, dist: torch.Tensor, sd, distilled_type: str = "mean"):
    """models/maest.py:905-925; head.0 is nn.LayerNorm with the default eps 1e-5 (:570-575)."""
    dt = cls.dtype
    feats = (cls + dist) / 2

    def _h(z):
        z = layer_norm(z, sd["head.0.weight"].to(dt), sd["head.0.bias"].to(dt), 1e-5)
        return z @ sd["head.1.weight"].to(dt).t() + sd["head.1.bias"].to(dt)

    if distilled_type == "mean":
        return _h(feats), feats
    lc = _h(cls)
    ld = dist @ sd["head_dist.weight"].to(dt).t() + sd["head_dist.bias"].to(dt)
    return lc, ld, feats


# ----------------------------------------------------------------------------------------------
# forward dispatch (models/maest.py:831-933) and predict_labels (:935-939)
# ----------------------------------------------------------------------------------------------
def to_mel_batch(x: torch.Tensor, img_t: int, melspectrogram_input: bool = False,
                 dtype=torch.float32) -> torch.Tensor:
    """Input-rank dispatch of MAEST.forward -> mel batch [B, 96, T]."""
    if x.dim() == 1:
        assert not melspectrogram_input
        m = logmel(x, dtype)                              # [96, T]
        if m.shape[1] >= img_t:
            trim = m.shape[1] % img_t
            if trim:
                m = m[:, :-trim]
            return m.reshape(N_MELS, -1, img_t).transpose(0, 1)
        return m[None]
    if x.dim() == 2 and melspectrogram_input:
        m = x.to(dtype)
        trim = m.shape[1] % img_t
        if trim:
            m = m[:, :-trim]
        return m.reshape(N_MELS, -1, img_t).transpose(0, 1)
    if x.dim() == 2:
        return logmel(x, dtype)
    if x.dim() == 3:
        return x.to(dtype)
    return x[:, 0].to(dtype)


def forward(x: torch.Tensor, sd, img_t: int, transformer_block: int = -1,
            return_self_attention: bool = False, melspectrogram_input: bool = False,
            distilled_type: str = "mean", dtype=torch.float32, t_offset: int = 0,
            keep_t=None, keep_f=None, keep_seq=None):
    mel = to_mel_batch(x, img_t, melspectrogram_input, dtype)
    tok = patch_tokens(mel, sd, t_offset=t_offset, keep_t=keep_t, keep_f=keep_f, keep_seq=keep_seq)
    out = encode(tok, sd, transformer_block, return_self_attention)
    if transformer_block != -1:
        return None, out
    return head(out[0], out[1], sd, distilled_type)


def predict_labels(x: torch.Tensor, sd, img_t: int, dtype=torch.float32) -> torch.Tensor:
    logits = forward(x, sd, img_t, dtype=dtype)[0]
    return torch.sigmoid(logits).mean(dim=0)


# ----------------------------------------------------------------------------------------------
# training step (models/module.py:73-102, helpers/mixup.py:5-12)
# ----------------------------------------------------------------------------------------------
def mixup(x: torch.Tensor, y: torch.Tensor, rn_indices: torch.Tensor, lam: torch.Tensor):
    B = x.shape[0]
    lam4 = lam.reshape(B, 1, 1, 1).to(x.dtype)
    lam2 = lam.reshape(B, 1).to(y.dtype)
    x = x * lam4 + x[rn_indices] * (1.0 - lam4)
    y = y * lam2 + y[rn_indices] * (1.0 - lam2)
    return x, y


def bce_with_logits(z: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    # mean over all elements of max(z,0) - z*y + log(1 + exp(-|z|))
    return (z.clamp(min=0) - z * y + torch.log1p(torch.exp(-z.abs()))).mean()


def training_loss(x: torch.Tensor, y: torch.Tensor, sd, rn_indices=None, lam=None, t_offset: int = 0,
                  keep_t=None, keep_f=None, keep_seq=None, dtype=torch.float32) -> torch.Tensor:
    """Loss of Module.training_step for mel batch x [B,1,96,T] and targets y [B,C]; host RNG draws
    (mixup permutation/lambda, time offset, patchout indices) are passed in."""
    x = x.to(dtype)
    y = y.to(dtype)
    if rn_indices is not None:
        x, y = mixup(x, y, rn_indices, lam)
    logits, _ = forward(x, sd, img_t=x.shape[-1], dtype=dtype, t_offset=t_offset, keep_t=keep_t,
                        keep_f=keep_f, keep_seq=keep_seq)
    return bce_with_logits(logits, y)
