"""Torch-facing wrappers over the C ABI (one function per entry point of include/maest_b200.h).

torch is used for device memory and streams only; all arithmetic happens in libmaest_b200.so.  Every wrapper
requires CUDA tensors and raises if the library is unavailable — there is no eager/CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib

F16, BF16, F32 = _lib.F16, _lib.BF16, _lib.F32
_TORCH2DT = {torch.float16: F16, torch.bfloat16: BF16, torch.float32: F32}
_DT2TORCH = {F16: torch.float16, BF16: torch.bfloat16, F32: torch.float32}


def op_dtype_code(dtype) -> int:
    if isinstance(dtype, int):
        return dtype
    if isinstance(dtype, str):
        dtype = {"fp16": torch.float16, "f16": torch.float16, "float16": torch.float16, "bf16": torch.bfloat16,
                 "bfloat16": torch.bfloat16}[dtype]
    return _TORCH2DT[dtype]


def _need_cuda(*ts: torch.Tensor):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("maest_b200 runs on CUDA (sm_100a) only; got a CPU tensor — there is no CPU fallback")


def h2d(t: torch.Tensor, device) -> torch.Tensor:
    """Host tensor -> device WITHOUT a host/device sync: staged through pinned memory, non-blocking (a plain `.to(device)` of a
    pageable tensor waits for everything already queued on the stream -- once per draw of mixup / patchout indices that made the
    training step host-bound: 45.7 ms of host time for 44.5 ms of kernels).  Device tensors pass through."""
    if t.device.type != "cpu":
        return t.to(device)
    return t.pin_memory().to(device, non_blocking=True)


def _lib_for(t: torch.Tensor):
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    return _lib.init(dev)


# A training step makes ~300 calls into the library; `torch.cuda.current_stream()` and the `torch.cuda.device(...)` guard cost
# ~15 + ~8 us of host time each.  Inside a `launch_scope(device)` both are resolved once (the step runs on one device and one
# stream) and every op reuses them; outside a scope each op resolves them itself as before.
import threading


class _Scope(threading.local):      # per thread: autograd runs the backward on its own worker thread
    def __init__(self):
        self.stream = None
        self.device = None


_SCOPE = _Scope()


class launch_scope:
    def __init__(self, device):
        self.device = torch.device(device)

    def __enter__(self):
        self._guard = torch.cuda.device(self.device)
        self._guard.__enter__()
        self._prev = (_SCOPE.stream, _SCOPE.device)
        _SCOPE.stream = torch.cuda.current_stream(self.device).cuda_stream
        _SCOPE.device = self.device
        return self

    def __exit__(self, *exc):
        _SCOPE.stream, _SCOPE.device = self._prev
        return self._guard.__exit__(*exc)


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL = _NullCtx()


def _dev(device):
    """Device guard for one library call (a no-op inside a launch_scope on the same device)."""
    return _NULL if _SCOPE.device == device else torch.cuda.device(device)


def _stream() -> int:
    s = _SCOPE.stream
    return s if s is not None else torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def set_gemm_mode(pair) -> None:
    """Forward GEMM tile shape: False/0 = one CTA per 128x256 tile, True/1 = a CTA pair (cta_group::2) per 256x256 tile,
    None/2 = per-shape choice (library default)."""
    mode = 2 if pair is None else int(pair)
    _lib.check(_lib.load().maest_set_gemm_mode(mode), "set_gemm_mode")


def logmel(wav: torch.Tensor) -> torch.Tensor:
    """[B, S] (or [S]) fp32 CUDA waveform -> [B, 96, T] (or [96, T]) normalised log-mel."""
    _need_cuda(wav)
    squeeze = wav.dim() == 1
    w = wav.reshape(1, -1) if squeeze else wav
    if w.dtype != torch.float32:
        w = w.float()
    if w.stride(-1) != 1:
        w = w.contiguous()
    B, S = w.shape
    T = 1 + S // 256
    mel = torch.empty((B, 96, T), device=w.device, dtype=torch.float32)
    with _dev(w.device):
        lib = _lib_for(w)
        _lib.check(lib.maest_logmel_fwd(w.data_ptr(), B, S, w.stride(0), mel.data_ptr(), _stream()), "logmel")
    return mel[0] if squeeze else mel


def logmel_raw16(wav: torch.Tensor, framing: str = "torchaudio") -> torch.Tensor:
    """[B, S] fp32 CUDA waveform -> [B, T, 96] fp16, un-normalised log10(1 + 1e4 mel), time-major: the layout of the
    reference's .mmap training files (helpers/melspectrogram_extractor.py).  framing "torchaudio": the model's own front-end
    (reflect padding, periodic Hann, T = 1 + S // 256); "essentia": the framing of the reference's offline extractor
    (FrameCutter centred frames with zero padding, symmetric Hann, T = ceil(S / 256))."""
    if framing not in ("torchaudio", "essentia"):
        raise ValueError(f"framing must be 'torchaudio' or 'essentia', got {framing!r}")
    ess = framing == "essentia"
    _need_cuda(wav)
    w = wav.reshape(1, -1) if wav.dim() == 1 else wav
    if w.dtype != torch.float32:
        w = w.float()
    if w.stride(-1) != 1:
        w = w.contiguous()
    B, S = w.shape
    out = torch.empty((B, (S + 255) // 256 if ess else 1 + S // 256, 96), device=w.device, dtype=torch.float16)
    with _dev(w.device):
        lib = _lib_for(w)
        _lib.check(lib.maest_logmel_raw16_fwd(w.data_ptr(), B, S, w.stride(0), out.data_ptr(), int(ess), _stream()), "logmel_raw16")
    return out[0] if wav.dim() == 1 else out


def ap_roc(y_true: torch.Tensor, y_score: torch.Tensor):
    """Per-class (average precision, ROC AUC) of [n, C] CUDA tensors with scikit-learn's semantics (ties share a threshold).
    The sort is torch's (library); the threshold scan runs in the ap_roc kernel.  Returns fp64 [C], fp64 [C], int32 [C] positives."""
    _need_cuda(y_true, y_score)
    assert y_true.shape == y_score.shape and y_score.dim() == 2
    n, C = y_score.shape
    s, idx = torch.sort(y_score.float(), dim=0, descending=True, stable=True)
    lab = torch.gather(y_true.float(), 0, idx).contiguous()
    s = s.contiguous()
    ap = torch.empty(C, device=s.device, dtype=torch.float64)
    auc = torch.empty(C, device=s.device, dtype=torch.float64)
    npos = torch.empty(C, device=s.device, dtype=torch.int32)
    with _dev(s.device):
        lib = _lib_for(s)
        _lib.check(lib.maest_ap_roc_fwd(s.data_ptr(), lab.data_ptr(), n, C, ap.data_ptr(), auc.data_ptr(), npos.data_ptr(), _stream()), "ap_roc")
    return ap, auc, npos


def mel_ingest(raw: torch.Tensor, frames_read: Optional[torch.Tensor] = None, roll_shift: Optional[torch.Tensor] = None,
               norm_mean: Optional[float] = None, norm_std: Optional[float] = None) -> torch.Tensor:
    """raw fp16 CUDA [B, T, 96] (file windows, time-major) -> [B, 1, 96, T] fp16: zero-pad centring, normalisation and time
    roll of the reference's loader (discogs/dataset.py:120-139, discogs/datamodule.py:111-137) in one kernel."""
    _need_cuda(raw, frames_read, roll_shift)
    assert raw.dtype == torch.float16 and raw.dim() == 3 and raw.shape[2] == 96 and raw.is_contiguous()
    B, T, _ = raw.shape
    for t in (frames_read, roll_shift):
        assert t is None or (t.dtype == torch.int32 and t.numel() == B and t.is_contiguous())
    do_norm = norm_mean is not None
    out = torch.empty((B, 1, 96, T), device=raw.device, dtype=torch.float16)
    with _dev(raw.device):
        lib = _lib_for(raw)
        _lib.check(lib.maest_mel_ingest_fwd(raw.data_ptr(), _p(frames_read), _p(roll_shift), B, T, int(do_norm),
                                            float(norm_mean or 0.0), float(norm_std or 1.0), out.data_ptr(), _stream()),
                   "mel_ingest")
    return out


def layernorm16(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float, op_dtype=F16,
                save_stats: bool = False, out: Optional[torch.Tensor] = None):
    _need_cuda(x, w, b)
    assert x.dtype == torch.float32 and x.shape[-1] == 768 and x.is_contiguous()
    rows = x.numel() // 768
    dt = op_dtype_code(op_dtype)
    y = out if out is not None else torch.empty(x.shape, device=x.device, dtype=_DT2TORCH[dt])
    assert y.dtype == _DT2TORCH[dt] and y.numel() == x.numel() and y.is_contiguous()
    mean = rstd = None
    if save_stats:
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    with _dev(x.device):
        lib = _lib_for(x)
        _lib.check(lib.maest_layernorm_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), dt, rows, float(eps),
                                           _p(mean), _p(rstd), _stream()), "layernorm")
    return (y, mean, rstd) if save_stats else y


def linear(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], epilogue: int,
           resid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
           addend: Optional[torch.Tensor] = None, rows_per_group: int = 0, group_stride: int = 0,
           row_offset: int = 0) -> torch.Tensor:
    """out = epilogue(a @ w.T + bias).  a [M,K], w [N,K] 16-bit (same dtype), K contiguous."""
    _need_cuda(a, w, bias, resid, out, addend)
    assert a.dtype == w.dtype and a.dtype in (torch.float16, torch.bfloat16)
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1] and a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    dt = _TORCH2DT[a.dtype]
    if out is None:
        odt = a.dtype if epilogue in (_lib.EPI_STORE16, _lib.EPI_GELU16) else torch.float32
        out = torch.empty((M, N), device=a.device, dtype=odt)
    if epilogue == _lib.EPI_RESID32 and resid is None:
        resid = out
    with _dev(a.device):
        lib = _lib_for(a)
        _lib.check(lib.maest_linear_fwd(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _p(bias), M, N, K, dt,
                                        epilogue, out.data_ptr(), out.stride(-2), _p(resid), _p(addend),
                                        rows_per_group, group_stride, row_offset, _stream()), "linear")
    return out


def ln_fold(w16: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: Optional[torch.Tensor]):
    """Weight-side half of the LayerNorm folding: (W gamma, bias + W beta) as fp32 [N] vectors from the 16-bit weight copy."""
    _need_cuda(w16, gamma, beta, bias)
    N, K = w16.shape
    assert w16.is_contiguous() and gamma.numel() == K and beta.numel() == K
    wg = torch.empty(N, device=w16.device, dtype=torch.float32)
    bf = torch.empty(N, device=w16.device, dtype=torch.float32)
    with _dev(w16.device):
        lib = _lib_for(w16)
        _lib.check(lib.maest_ln_fold(w16.data_ptr(), gamma.data_ptr(), beta.data_ptr(), _p(bias), N, K, _TORCH2DT[w16.dtype],
                                     wg.data_ptr(), bf.data_ptr(), _stream()), "ln_fold")
    return wg, bf


def linear_ln(a: torch.Tensor, w16: torch.Tensor, bias: torch.Tensor, epilogue: int, ln_stats: torch.Tensor, ln_vec: torch.Tensor,
              out: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None, out16b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Linear with the neighbouring LayerNorm folded in (include/maest_b200.h: maest_linear_ln_fwd)."""
    _need_cuda(a, w16, bias, ln_stats, ln_vec, out, resid, out16b)
    M, K = a.shape
    N = w16.shape[0]
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32 if epilogue == _lib.EPI_RESID32_LN else a.dtype)
    with _dev(a.device):
        lib = _lib_for(a)
        _lib.check(lib.maest_linear_ln_fwd(a.data_ptr(), a.stride(0), w16.data_ptr(), w16.stride(0), _p(bias), M, N, K, _TORCH2DT[a.dtype],
                                           epilogue, out.data_ptr(), out.stride(0), _p(resid), ln_stats.data_ptr(), ln_vec.data_ptr(),
                                           _p(out16b), _stream()), "linear_ln")
    return out


def ln_finalize(partials: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """partials fp32 [n/32, M, 4] from a RESID32_LN producer -> stats fp32 [M, 2] = (rstd, -mean * rstd)."""
    _need_cuda(partials)
    nparts, M, _ = partials.shape
    stats = torch.empty((M, 2), device=partials.device, dtype=torch.float32)
    with _dev(partials.device):
        lib = _lib_for(partials)
        _lib.check(lib.maest_ln_finalize(partials.data_ptr(), M, nparts * 128, float(eps), stats.data_ptr(), _stream()), "ln_finalize")
    return stats


def attention(qkv: torch.Tensor, B: int, N: int, heads: int = 12, variant: int = 8, save_lse: bool = False,
              out: Optional[torch.Tensor] = None):
    """qkv [B*N, 3*heads*64] 16-bit -> o [B*N, heads*64] 16-bit (and, for training, lse fp32 [B, heads, N]).
    variant 8 (default): chains kernel, 3 x 128 keys + epilogue warpgroup (csrc/attention_chain.cuh); 3: the same without the
    epilogue warpgroup; 4: 4 x 96 keys; 5: split columns; 7: lean ring; 9: 8 + early PV;
    0 / 1 / 2: the round-1 kernels (profiles/r02_attention_notes.md has the A/B of all of them)."""
    _need_cuda(qkv)
    assert qkv.is_contiguous() and qkv.shape == (B * N, 3 * heads * 64)
    if out is None:
        out = torch.empty((B * N, heads * 64), device=qkv.device, dtype=qkv.dtype)
    assert out.is_contiguous() and out.shape == (B * N, heads * 64) and out.dtype == qkv.dtype
    lse = torch.empty((B, heads, N), device=qkv.device, dtype=torch.float32) if save_lse else None
    with _dev(qkv.device):
        lib = _lib_for(qkv)
        _lib.check(lib.maest_attention_fwd(qkv.data_ptr(), out.data_ptr(), _p(lse), B, N, heads, _TORCH2DT[qkv.dtype], variant,
                                           _stream()), "attention")
    return (out, lse) if save_lse else out


def attention_bwd(qkv: torch.Tensor, o: torch.Tensor, d_o: torch.Tensor, lse: torch.Tensor, B: int, N: int, heads: int = 12,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Gradient of `attention` w.r.t. the packed qkv activation: returns dqkv [B*N, 3*heads*64] 16-bit."""
    _need_cuda(qkv, o, d_o, lse)
    assert qkv.is_contiguous() and o.is_contiguous() and d_o.is_contiguous() and lse.is_contiguous()
    assert o.dtype == qkv.dtype == d_o.dtype
    dev = qkv.device
    dqkv = out if out is not None else torch.empty_like(qkv)
    delta = torch.empty((B, heads, N), device=dev, dtype=torch.float32)
    dq32 = torch.empty((B * N, heads * 64), device=dev, dtype=torch.float32)
    with _dev(dev):
        lib = _lib_for(qkv)
        _lib.check(lib.maest_attention_bwd(qkv.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr(), delta.data_ptr(),
                                           dq32.data_ptr(), dqkv.data_ptr(), B, N, heads, _TORCH2DT[qkv.dtype], _stream()),
                   "attention_bwd")
    return dqkv


def gemm(a: torch.Tensor, b: torch.Tensor, epilogue: int, M: int, N: int, K: int, a_mn: bool = False, b_mn: bool = False,
         bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
         aux16: Optional[torch.Tensor] = None, k_splits: int = 1, colsum_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = epilogue(sum_k A(m,k) B(n,k)); a_mn / b_mn: operand stored [K, M] / [K, N] (see include/maest_b200.h).
    colsum_out (EPI_GELUBWD16 only): fp32 [N], the column sums of `out` are accumulated into it (the bias gradient)."""
    _need_cuda(a, b, bias, out, resid, aux16, colsum_out)
    assert colsum_out is None or (epilogue == _lib.EPI_GELUBWD16 and colsum_out.dtype == torch.float32 and colsum_out.numel() == N)
    assert a.dtype == b.dtype and a.dtype in (torch.float16, torch.bfloat16) and a.stride(-1) == 1 and b.stride(-1) == 1
    if out is None:
        odt = a.dtype if epilogue in (_lib.EPI_STORE16, _lib.EPI_GELU16, _lib.EPI_GELUBWD16) else torch.float32
        out = torch.empty((M, N), device=a.device, dtype=odt)
    if epilogue == _lib.EPI_RESID32 and resid is None:
        resid = out
    with _dev(a.device):
        lib = _lib_for(a)
        _lib.check(lib.maest_gemm(a.data_ptr(), a.stride(0), int(a_mn), b.data_ptr(), b.stride(0), int(b_mn), _p(bias), M, N, K,
                                  _TORCH2DT[a.dtype], epilogue, out.data_ptr(), out.stride(-2), _p(resid), _p(colsum_out), 0, 0, 0,
                                  _p(aux16), int(k_splits), _stream()), "gemm")
    return out


def wgrad_splits(M: int, N: int, K: int, sms: int = 148) -> int:
    """K splits of a weight-gradient GEMM ([M, N] output, reduction over K tokens) on the persistent 1-CTA kernel: the work items
    (tiles x splits) should fill whole waves of `sms` CTAs.  Minimises  waves x (k-blocks per item + a fixed per-item cost of
    ~8 k-blocks for pipeline fill and the atomic epilogue).  The old rule, ceil(2 sms / tiles), left 19-31 % of the last wave idle
    (proj: 18 tiles x 17 splits = 306 items = 2.07 waves)."""
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    kb = (K + 63) // 64
    best, best_cost = 1, None
    for s in range(1, max(1, min(64, kb // 4)) + 1):
        items = tiles * s
        waves = (items + sms - 1) // sms
        cost = waves * ((kb + s - 1) // s + 8)
        if best_cost is None or cost < best_cost:
            best, best_cost = s, cost
    return best


def mixup(x: torch.Tensor, perm: torch.Tensor, lam: torch.Tensor) -> torch.Tensor:
    """x [B, ...] (fp16|fp32) -> fp32 blend x*lam + x[perm]*(1-lam) (models/module.py:77-86)."""
    _need_cuda(x, perm, lam)
    x = x.contiguous()
    B = x.shape[0]
    L = x.numel() // B
    out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
    with _dev(x.device):
        lib = _lib_for(x)
        _lib.check(lib.maest_mixup_fwd(x.data_ptr(), _TORCH2DT[x.dtype], perm.to(torch.int32).contiguous().data_ptr(),
                                       lam.float().contiguous().data_ptr(), out.data_ptr(), B, L, _stream()), "mixup")
    return out


def bce_logits(logits: torch.Tensor, targets: torch.Tensor):
    """(loss scalar, dlogits) of F.binary_cross_entropy_with_logits (mean)."""
    _need_cuda(logits, targets)
    z = logits.float().contiguous()
    y = targets.float().contiguous()
    loss = torch.empty((), device=z.device, dtype=torch.float32)
    dz = torch.empty_like(z)
    with _dev(z.device):
        lib = _lib_for(z)
        _lib.check(lib.maest_bce_logits_fwd(z.data_ptr(), y.data_ptr(), z.numel(), loss.data_ptr(), dz.data_ptr(), _stream()), "bce")
    return loss, dz


def layernorm_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, op_dtype, dx16: Optional[torch.Tensor] = None,
                  dx_colsum: Optional[torch.Tensor] = None):
    """dx += LN'(dy) (+ dgamma, dbeta); dx_colsum[768] += column sums of the updated dx (bias gradient of the layer upstream)."""
    _need_cuda(dy, x, dx)
    rows = x.numel() // 768
    with _dev(x.device):
        lib = _lib_for(x)
        _lib.check(lib.maest_layernorm_bwd(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                           dx.data_ptr(), _p(dx16), op_dtype_code(op_dtype), dgamma.data_ptr(), dbeta.data_ptr(),
                                           _p(dx_colsum), rows, _stream()), "layernorm_bwd")


def colsum(x: torch.Tensor, out: torch.Tensor):
    _need_cuda(x, out)
    M, N = x.shape
    with _dev(x.device):
        lib = _lib_for(x)
        _lib.check(lib.maest_colsum(x.data_ptr(), _TORCH2DT[x.dtype], x.stride(0), M, N, out.data_ptr(), _stream()), "colsum")


def cast_rows16(src: torch.Tensor, rows: int, op_dtype, rows_per_group: int = 0, group_stride: int = 0, row_offset: int = 0,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(src)
    dt = op_dtype_code(op_dtype)
    if out is None:
        out = torch.empty((rows, 768), device=src.device, dtype=_DT2TORCH[dt])
    with _dev(src.device):
        lib = _lib_for(src)
        _lib.check(lib.maest_cast_rows16(src.data_ptr(), out.data_ptr(), out.stride(0), rows, rows_per_group, group_stride,
                                         row_offset, dt, _stream()), "cast_rows16")
    return out


def cast16(src: torch.Tensor, op_dtype=F16) -> torch.Tensor:
    _need_cuda(src)
    s = src.detach().float().contiguous()
    dt = op_dtype_code(op_dtype)
    dst = torch.empty(s.shape, device=s.device, dtype=_DT2TORCH[dt])
    with _dev(s.device):
        lib = _lib_for(s)
        _lib.check(lib.maest_cast_to16(s.data_ptr(), dst.data_ptr(), s.numel(), dt, _stream()), "cast16")
    return dst


def keep_ft_tensor(keep_f: Optional[Sequence[int]], keep_t: Optional[Sequence[int]], Fp: int, Tp: int,
                   keep_seq: Optional[Sequence[int]], device) -> Optional[torch.Tensor]:
    """Kept patch-grid cells in sequence order as int32 (f << 16 | t); None when nothing is dropped."""
    if keep_f is None and keep_t is None and keep_seq is None:
        return None
    f = torch.arange(Fp) if keep_f is None else torch.as_tensor(list(keep_f), dtype=torch.long)
    t = torch.arange(Tp) if keep_t is None else torch.as_tensor(list(keep_t), dtype=torch.long)
    ft = (f[:, None] * 65536 + t[None, :]).reshape(-1)
    if keep_seq is not None:
        ft = ft[torch.as_tensor(list(keep_seq), dtype=torch.long)]
    return h2d(ft.to(torch.int32), device)


def patch_tokens(mel: torch.Tensor, w_pe16: torch.Tensor, conv_bias, freq_pe, time_pe, cls_token, dist_token,
                 new_pos_embed, keep_ft: Optional[torch.Tensor] = None, t_offset: int = 0, return_patches: bool = False):
    """mel [B,96,T] (fp32|fp16) -> tokens fp32 [B, 2+P, 768].  Parameter tensors in the reference's shapes."""
    _need_cuda(mel, w_pe16)
    assert mel.dim() == 3 and mel.is_contiguous() and mel.dtype in (torch.float32, torch.float16)
    B, Fm, T = mel.shape
    Fp = (Fm - 16) // 10 + 1
    Tp = (T - 16) // 10 + 1
    Wt = time_pe.shape[-1]
    P = Fp * Tp if keep_ft is None else int(keep_ft.numel())
    tokens = torch.empty((B, 2 + P, 768), device=mel.device, dtype=torch.float32)
    with _dev(mel.device):
        lib = _lib_for(mel)
        ws_bytes = lib.maest_patch_workspace_bytes(B, P)
        ws = torch.empty(ws_bytes, device=mel.device, dtype=torch.uint8)
        _lib.check(lib.maest_patch_tokens_fwd(
            mel.data_ptr(), _TORCH2DT[mel.dtype], B, T, w_pe16.data_ptr(), _TORCH2DT[w_pe16.dtype],
            conv_bias.data_ptr(), freq_pe.data_ptr(), Fp, time_pe.data_ptr(), Wt, cls_token.data_ptr(),
            dist_token.data_ptr(), new_pos_embed.data_ptr(), _p(keep_ft), P, int(t_offset), tokens.data_ptr(),
            ws.data_ptr(), ws_bytes, _stream()), "patch_tokens")
    if return_patches:   # the gathered patch operand [B*P, 256] (op16) is the B operand of the conv weight gradient
        a16 = ws[: B * P * 512].view(w_pe16.dtype).view(B * P, 256)
        return tokens, a16
    return tokens


def wave_tokens(wav: torch.Tensor, w_pe16: torch.Tensor, conv_bias, freq_pe, time_pe, cls_token, dist_token, new_pos_embed,
                keep_ft: Optional[torch.Tensor] = None, t_offset: int = 0) -> torch.Tensor:
    """K1 + K2 in ONE C-ABI call: waveform [B, S] fp32 -> tokens fp32 [B, 2+P, 768] (the mel stays in the call's workspace)."""
    _need_cuda(wav, w_pe16)
    assert wav.dim() == 2 and wav.dtype == torch.float32 and wav.stride(-1) == 1
    B, S = wav.shape
    T = 1 + S // 256
    Fp, Tp = 9, (T - 16) // 10 + 1
    Wt = time_pe.shape[-1]
    P = Fp * Tp if keep_ft is None else int(keep_ft.numel())
    tokens = torch.empty((B, 2 + P, 768), device=wav.device, dtype=torch.float32)
    with _dev(wav.device):
        lib = _lib_for(wav)
        ws_bytes = lib.maest_wave_tokens_workspace_bytes(B, S, P)
        ws = torch.empty(ws_bytes, device=wav.device, dtype=torch.uint8)
        _lib.check(lib.maest_wave_tokens_fwd(
            wav.data_ptr(), B, S, wav.stride(0), w_pe16.data_ptr(), _TORCH2DT[w_pe16.dtype], conv_bias.data_ptr(), freq_pe.data_ptr(),
            Fp, time_pe.data_ptr(), Wt, cls_token.data_ptr(), dist_token.data_ptr(), new_pos_embed.data_ptr(), _p(keep_ft), P,
            int(t_offset), tokens.data_ptr(), ws.data_ptr(), ws_bytes, _stream()), "wave_tokens")
    return tokens


def encoder(x: torch.Tensor, B: int, N: int, block_table, n_blocks: int, last_attn_only: bool, op_dtype,
            attn_variant: int = 8, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Run n_blocks transformer blocks in place on the fp32 residual stream x [B*N, 768]."""
    _need_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous()
    with _dev(x.device):
        lib = _lib_for(x)
        need = lib.maest_encoder_workspace_bytes(B * N)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, device=x.device, dtype=torch.uint8)
        _lib.check(lib.maest_encoder_fwd(x.data_ptr(), B, N, block_table, n_blocks, int(bool(last_attn_only)),
                                         op_dtype_code(op_dtype), attn_variant, workspace.data_ptr(),
                                         workspace.numel(), _stream()), "encoder")
    return x


def pool_head(x: torch.Tensor, B: int, N: int, norm_w, norm_b, hln_w, hln_b, head_w, head_b, hdist_w=None,
              hdist_b=None, separated: bool = False, save_ln: bool = False):
    _need_cuda(x)
    C_ = head_w.shape[0]
    dev = x.device
    logits = torch.empty((B, C_), device=dev, dtype=torch.float32)
    logits_dist = torch.empty((B, C_), device=dev, dtype=torch.float32) if separated else None
    feats = torch.empty((B, 768), device=dev, dtype=torch.float32)
    ln_cls = torch.empty((B, 768), device=dev, dtype=torch.float32) if save_ln else None
    ln_dist = torch.empty((B, 768), device=dev, dtype=torch.float32) if save_ln else None
    with _dev(dev):
        lib = _lib_for(x)
        _lib.check(lib.maest_pool_head_fwd(x.data_ptr(), B, N, norm_w.data_ptr(), norm_b.data_ptr(), hln_w.data_ptr(),
                                           hln_b.data_ptr(), head_w.data_ptr(), head_b.data_ptr(), _p(hdist_w),
                                           _p(hdist_b), C_, _lib.HEAD_SEPARATED if separated else _lib.HEAD_MEAN,
                                           logits.data_ptr(), _p(logits_dist), feats.data_ptr(), _p(ln_cls),
                                           _p(ln_dist), _stream()), "pool_head")
    if save_ln:
        return logits, logits_dist, feats, ln_cls, ln_dist
    return logits, logits_dist, feats


def block_embedding(x: torch.Tensor, B: int, N: int) -> torch.Tensor:
    _need_cuda(x)
    emb = torch.empty((B, 3 * 768), device=x.device, dtype=torch.float32)
    with _dev(x.device):
        lib = _lib_for(x)
        _lib.check(lib.maest_block_embedding_fwd(x.data_ptr(), B, N, emb.data_ptr(), _stream()), "block_embedding")
    return emb
