"""GPU parity tests (run on the B200 box: `pytest -m gpu`).  Everything goes through the C ABI (ops.py -> libmaest_b200.so).

Checked against (a) the committed golden outputs of the unmodified reference (tests/golden/*.npz), (b) the CPU
oracle on the same seeded inputs, (c) size-independent properties at BASELINE.json's full sizes.

Tolerances (stated per BASELINE.json north_star: logits/embeddings <= 1e-3 relative):
  * log-mel (fp32 kernel)                      max-abs <= 2e-5 vs float64 oracle
  * fp16-operand path, logits / embeddings     rel-L2 <= 1e-3  vs the fp32 reference (same fp16/bf16-representable weights)
  * bf16-operand path                          rel-L2 <= 6e-3  (the reference's own autocast(bf16) is 6.4e-3 off its fp32 run)
"""
import numpy as np
import pytest
import torch

from maest_b200 import _lib, get_maest, ops, synth
from oracle import maest_oracle as O

pytestmark = pytest.mark.gpu

TOL_F16 = 1e-3
TOL_BF16 = 6e-3


def rel(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().flatten().cpu()
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def make(arch, grid_t, n_classes=400, **kw):
    m = get_maest(arch=arch, pretrained=False, n_classes=n_classes, **kw)
    m.load_state_dict(synth.synth_state_dict(grid_t, m.num_classes, seed=0), strict=False)
    return m.cuda().eval()


@pytest.fixture(scope="module")
def m10():
    return make("discogs-maest-10s-pw-129e", 62)


@pytest.fixture(scope="module")
def m30():
    return make("discogs-maest-30s-pw-129e", 187)


def test_native_library_is_loaded():
    lib = _lib.init(0)
    assert lib.maest_abi_version() == _lib.ABI_VERSION
    maps = open("/proc/self/maps").read()
    assert "libmaest_b200.so" in maps


# ------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("case", ["waveA_2x10s", "waveB_10s", "odd_length", "waveA_30s", "min_length"])
def test_logmel_vs_oracle(case):
    x = {"waveA_2x10s": synth.wave_a(2, 160000), "waveB_10s": synth.wave_b(160000)[None],
         "odd_length": synth.wave_a(3, 48123, seed=3), "waveA_30s": synth.wave_a(1, 480000),
         "min_length": synth.wave_a(2, 257, seed=5)}[case]
    ref = O.logmel(x, torch.float64)
    got = ops.logmel(x.cuda()).cpu()
    assert got.shape == ref.shape
    assert float((got.double() - ref).abs().max()) < 2e-5


def test_logmel_vs_reference_golden(golden):
    g = golden["c2"]
    got = ops.logmel(synth.wave_a(2, 160000).cuda()).cpu().numpy()
    assert np.abs(got[0] - g["mel0"]).max() < 5e-5
    assert np.abs(got[1][:, ::5] - g["mel1_sub"]).max() < 5e-5
    gb = ops.logmel(synth.wave_b(160000).cuda()).cpu().numpy()
    assert np.abs(gb - g["mel_waveb"]).max() < 5e-5


def test_logmel_rejects_too_short():
    with pytest.raises(RuntimeError, match="too short"):
        ops.logmel(torch.rand(1, 256).cuda())


# ------------------------------------------------------------------------------------------ GEMM / attention units
@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("shape", [(300, 256, 64), (1000, 2304, 768), (257, 768, 3072), (1674, 768, 256), (129, 3072, 768)])
def test_linear_epilogues_vs_fp64(dt, shape):
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).to(dt).cuda()
    W = (torch.randn(N, K, generator=g) * 0.05).to(dt).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = A.double() @ W.double().t() + bias.double()
    tol16 = 1e-3 if dt == torch.float16 else 5e-3
    assert rel(ops.linear(A, W, bias, _lib.EPI_STORE32), ref) < 1e-5
    assert rel(ops.linear(A, W, bias, _lib.EPI_STORE16), ref) < tol16
    assert rel(ops.linear(A, W, bias, _lib.EPI_GELU16), torch.nn.functional.gelu(ref)) < tol16
    x0 = torch.randn(M, N, generator=g).cuda()
    out = x0.clone()
    ops.linear(A, W, bias, _lib.EPI_RESID32, resid=out, out=out)          # in place: the add is a TMA reduce into out
    assert rel(out, x0.double() + ref) < 1e-5
    out2 = torch.empty_like(x0)
    ops.linear(A, W, bias, _lib.EPI_RESID32, resid=x0, out=out2)           # out of place: load-add-store epilogue
    assert rel(out2, x0.double() + ref) < 1e-5
    assert float((out2 - out).abs().max()) <= 4e-6 * float(out.abs().max())     # the two forms differ by the order of two additions
    again = x0.clone()
    ops.linear(A, W, bias, _lib.EPI_RESID32, resid=again, out=again)
    assert torch.equal(again, out)                                          # one add per element: deterministic


def test_linear_identity_is_bit_exact():
    A = torch.randn(256, 256).half().cuda()
    C = ops.linear(A, torch.eye(256).half().cuda(), None, _lib.EPI_STORE32)
    assert torch.equal(C, A.float())


def test_linear_row_remap_and_addend():
    g = torch.Generator().manual_seed(3)
    B, P, N, K = 6, 50, 768, 256
    A = (torch.randn(B * P, K, generator=g) * 0.5).half().cuda()
    W = (torch.randn(N, K, generator=g) * 0.05).half().cuda()
    add = torch.randn(P, N, generator=g).cuda()
    out = torch.zeros(B, 2 + P, N).cuda()
    ops.linear(A, W, None, _lib.EPI_STORE32, out=out.view(-1, N), addend=add, rows_per_group=P, group_stride=2 + P, row_offset=2)
    exp = (A.double() @ W.double().t()).view(B, P, N) + add.double()
    assert rel(out[:, 2:], exp) < 1e-5 and bool((out[:, :2] == 0).all())


def test_gelu_epilogue_matches_exact_erf_gelu():
    # A = x on the diagonal trick: out = gelu(x * 1) for a sweep of x, vs torch's erf GELU in float64
    x = torch.linspace(-8, 8, 256 * 128).reshape(128, 256)
    A = torch.zeros(128, 256)
    A[:, 0] = 1.0
    W = torch.zeros(256, 256)
    out = ops.linear(A.half().cuda(), W.half().cuda(), None, _lib.EPI_GELU16).float().cpu()
    assert bool((out == 0).all())                       # gelu(0) == 0 exactly
    bias_sweep = torch.linspace(-8, 8, 256)
    out = ops.linear(A.half().cuda(), W.half().cuda(), bias_sweep.cuda(), _lib.EPI_GELU16).float().cpu()
    ref = torch.nn.functional.gelu(bias_sweep.double()).float()
    assert float((out[0] - ref).abs().max()) < 5e-3 and rel(out[0], ref) < 5e-4


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("BN", [(1, 128), (2, 100), (2, 560), (1, 1685), (3, 866), (1, 3), (2, 129), (1, 257), (5, 200)])
@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 7, 8, 9])
def test_attention_vs_fp64(dt, BN, variant):
    B, N = BN
    g = torch.Generator().manual_seed(B * 1000 + N)
    qkv = torch.randn(B * N, 2304, generator=g).to(dt).cuda()
    q, k, v = qkv.view(B, N, 3, 12, 64).permute(2, 0, 3, 1, 4).double()
    ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, -1) @ v).transpose(1, 2).reshape(B * N, 768)
    o = ops.attention(qkv, B, N, 12, variant)
    assert not torch.isnan(o.float()).any()
    assert rel(o, ref) < (1e-3 if dt == torch.float16 else 6e-3)


def test_attention_sharp_scores_and_rescale_path():
    g = torch.Generator().manual_seed(11)
    B, N = 1, 700
    qkv = (torch.randn(B * N, 2304, generator=g) * 4).half().cuda()
    # make later keys much larger than earlier ones so the running max jumps by > 2^8 between KV tiles
    qkv.view(N, 3, 12, 64)[400:, 1] *= 3
    q, k, v = qkv.view(B, N, 3, 12, 64).permute(2, 0, 3, 1, 4).double()
    ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, -1) @ v).transpose(1, 2).reshape(B * N, 768)
    for variant in (0, 1, 2, 3, 4, 5, 7, 8, 9):
        assert rel(ops.attention(qkv, B, N, 12, variant), ref) < 1e-3


@pytest.mark.parametrize("variant", [3, 4, 5, 7, 8, 9])
def test_attention_chain_kernel_lse_and_exact_redo(variant):
    """The chains kernel (default): log-sum-exp for the backward pass, and rows whose scores leave the fixed reference's range
    (here: a few query rows scaled up so that later keys beat the first KV tile by far more than 2^16) take the exact redo."""
    g = torch.Generator().manual_seed(5)
    B, N = 2, 700
    qkv = torch.randn(B * N, 2304, generator=g).half().cuda()
    qkv.view(B * N, 3, 12, 64)[300:310, 0] *= 40          # ten sharp query rows per ... (rows 300-309 of clip 0)
    qkv.view(B * N, 3, 12, 64)[1000:1003, 0] *= 60
    q, k, v = qkv.view(B, N, 3, 12, 64).permute(2, 0, 3, 1, 4).double()
    s = (q @ k.transpose(-1, -2)) * 0.125
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * N, 768)
    lse_ref = torch.logsumexp(s, -1) * 1.4426950408889634
    o, lse = ops.attention(qkv, B, N, 12, variant, save_lse=True)
    assert not torch.isnan(o.float()).any()
    assert rel(o, ref) < 1e-3
    assert float((lse.double() - lse_ref).abs().max()) < 2e-3
    o2, _ = ops.attention(qkv, B, N, 12, variant, save_lse=True)
    assert torch.equal(o, o2)


def test_layernorm_head_embedding_units():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1000, 768, generator=g) * 3 + 0.5
    w, b = torch.randn(768, generator=g), torch.randn(768, generator=g)
    ref = torch.nn.functional.layer_norm(x.double(), (768,), w.double(), b.double(), 1e-6)
    assert rel(ops.layernorm16(x.cuda(), w.cuda(), b.cuda(), 1e-6, "fp16"), ref) < 5e-4
    assert rel(ops.layernorm16(x.cuda(), w.cuda(), b.cuda(), 1e-6, "bf16"), ref) < 4e-3
    xs = torch.randn(3, 77, 768, generator=g)
    emb = ops.block_embedding(xs.cuda(), 3, 77).cpu()
    assert rel(emb, torch.cat([xs[:, 0], xs[:, 1], xs[:, 2:].double().mean(1).float()], 1)) < 1e-6


# ------------------------------------------------------------------------------------------ K2
def test_tokens_vs_oracle_and_golden(m10, golden):
    g = golden["c2"]
    sd = synth.synth_state_dict(62)
    mel = O.logmel(synth.wave_a(2, 160000), torch.float32).contiguous()
    got = m10.tokens_from_mel(mel.cuda()).cpu()
    ref = O.patch_tokens(mel.double(), sd)
    assert got.shape == (2, 560, 768)
    assert rel(got, ref) < 3e-4                          # mel patches are rounded to fp16 operands
    assert rel(got[:, :2], ref[:, :2]) < 1e-6            # cls / dist rows are pure fp32
    assert rel(got[:, list(g["row_probe"])], g["tokens_probe"]) < 3e-4


def test_tokens_with_patchout_vs_oracle(m10):
    sd = synth.synth_state_dict(62)
    mel = O.logmel(synth.wave_a(2, 160000), torch.float32).contiguous()
    keep_t, keep_f = list(range(0, 62, 3)), [0, 2, 3, 7]
    keep_seq = sorted(torch.randperm(len(keep_t) * len(keep_f), generator=torch.Generator().manual_seed(1))[:50].tolist())
    for ks in (None, keep_seq):
        kft = ops.keep_ft_tensor(keep_f, keep_t, 9, 62, ks, "cuda")
        got = ops.patch_tokens(mel.cuda(), m10._weight16("patch_embed.proj", m10.patch_embed.proj.weight),
                               m10.patch_embed.proj.bias.detach(), m10.freq_new_pos_embed.detach().reshape(768, -1).contiguous(),
                               m10.time_new_pos_embed.detach().reshape(768, -1).contiguous(), m10.cls_token.detach().reshape(-1),
                               m10.dist_token.detach().reshape(-1), m10.new_pos_embed.detach().reshape(2, 768).contiguous(),
                               keep_ft=kft, t_offset=0).cpu()
        ref = O.patch_tokens(mel.double(), sd, keep_t=keep_t, keep_f=keep_f, keep_seq=ks)
        assert got.shape == ref.shape and rel(got, ref) < 3e-4


def test_fp16_mel_input(m10):
    mel = (0.5 * torch.randn(2, 96, 625, generator=torch.Generator().manual_seed(7))).half()
    got = m10.tokens_from_mel(mel.cuda()).cpu()
    ref = O.patch_tokens(mel.double(), synth.synth_state_dict(62))
    assert rel(got, ref) < 3e-4


# ------------------------------------------------------------------------------------------ end to end vs the reference's golden outputs
def test_config2_logits_embeddings_vs_reference(m10, golden):
    g = golden["c2"]
    x = synth.wave_a(2, 160000).cuda()
    with torch.no_grad():
        lo, em = m10(x)
        assert rel(lo, g["logits"]) < TOL_F16 and rel(em, g["emb"]) < TOL_F16
        for k in (0, 6, 11):
            none, e = m10(x, transformer_block=k)
            assert none is None and rel(e, g[f"emb_block{k}"]) < TOL_F16
        e = m10(x, transformer_block=6, return_self_attention=True)[1]
        assert rel(e, g["emb_block6_selfattn"]) < TOL_F16


def test_forward_features_return_convention(m10):
    # models/maest.py:804-810: forward_features(-1) returns the final-LayerNorm'ed (cls, dist) rows; forward's features = their mean
    x = synth.wave_a(2, 160000).cuda()
    with torch.no_grad():
        mel = ops.logmel(x)
        cls, dist = m10.forward_features(mel[:, None])
        lo, feats = m10(x)
        e6 = m10.forward_features(mel, transformer_block=6)
    assert cls.shape == dist.shape == (2, 768) and e6.shape == (2, 2304)
    assert rel((cls + dist) / 2, feats) < 1e-6      # also: two runs of the folded-LayerNorm path are bit-reproducible


def test_block_by_block_drift_vs_reference(m10, golden):
    g = golden["c2"]
    rows = list(g["row_probe"])
    with torch.no_grad():
        tok = m10.tokens_from_mel(ops.logmel(synth.wave_a(2, 160000).cuda()))
        B, N, _ = tok.shape
        table = m10._blocks_ctypes()
        for i in range(12):
            one = (_lib.MaestBlockWeights * 1)(table[i])
            ops.encoder(tok.view(B * N, 768), B, N, one, 1, False, "fp16", 0)
            assert rel(tok[:, rows], g[f"block{i}_probe"]) < TOL_F16, i


def test_bf16_operands_vs_reference(golden):
    g = golden["c2"]
    m = make("discogs-maest-10s-pw-129e", 62, op_dtype="bf16")
    with torch.no_grad():
        lo, em = m(synth.wave_a(2, 160000).cuda())
    assert rel(lo, g["logits"]) < TOL_BF16 and rel(em, g["emb"]) < TOL_BF16


def test_layernorm_folding_units():
    """Producer / consumer epilogues of the LayerNorm folding against float64 math (models/maest.py:418-419 with :395,:405)."""
    g = torch.Generator().manual_seed(21)
    M, D, N2 = 300, 768, 2304
    x0 = (torch.randn(M, D, generator=g) * 2 + 0.7).cuda()                 # residual stream with a non-zero row mean
    a = (torch.randn(M, D, generator=g) * 0.5).half().cuda()
    wp = (torch.randn(D, D, generator=g) * 0.04).half().cuda()
    bp = torch.randn(D, generator=g).cuda() * 0.1
    gamma, beta = (1 + 0.3 * torch.randn(D, generator=g)).cuda(), (0.2 * torch.randn(D, generator=g)).cuda()
    w2 = (torch.randn(N2, D, generator=g) * 0.04).half().cuda()
    b2 = torch.randn(N2, generator=g).cuda() * 0.1
    parts = torch.full((D // 128, M, 4), float("nan"), device="cuda")
    h16 = torch.empty(M, D, device="cuda", dtype=torch.float16)
    x1 = ops.linear_ln(a, wp, bp, _lib.EPI_RESID32_LN, parts, gamma, resid=x0, out16b=h16)
    ref_x1 = x0.double() + a.double() @ wp.double().t() + bp.double()
    assert rel(x1, ref_x1) < 1e-6
    stats = ops.ln_finalize(parts, 1e-6)
    rstd = 1.0 / torch.sqrt(ref_x1.var(1, unbiased=False) + 1e-6)
    assert rel(stats[:, 0], rstd) < 1e-6 and rel(stats[:, 1], -ref_x1.mean(1) * rstd) < 1e-5
    assert rel(h16, ref_x1 * gamma.double()) < 4e-4
    # a large common offset of the rows must not cost precision (two-pass statistics inside each chunk)
    xb = x0 + 1000.0
    xb1 = ops.linear_ln(a, wp, bp, _lib.EPI_RESID32_LN, parts, gamma, resid=xb, out16b=torch.empty_like(h16))
    sb = ops.ln_finalize(parts, 1e-6)
    assert rel(sb[:, 0], 1.0 / torch.sqrt(xb1.double().var(1, unbiased=False) + 1e-6)) < 1e-5
    wg, bf = ops.ln_fold(w2, gamma, beta, b2)
    assert rel(wg, w2.double() @ gamma.double()) < 1e-6 and rel(bf, b2.double() + w2.double() @ beta.double()) < 1e-6
    ref_ln = torch.nn.functional.layer_norm(ref_x1, (D,), gamma.double(), beta.double(), 1e-6)
    y = ops.linear_ln(h16, w2, bf, _lib.EPI_STORE16_LN, stats, wg)
    assert rel(y, ref_ln @ w2.double().t() + b2.double()) < 1e-3
    y = ops.linear_ln(h16, w2, bf, _lib.EPI_GELU16_LN, stats, wg)
    assert rel(y, torch.nn.functional.gelu(ref_ln @ w2.double().t() + b2.double())) < 1e-3


def test_folded_layernorm_path_vs_reference(golden):
    """fuse_ln=True moves 23 of the 24 LayerNorms into the GEMM epilogues; it must match the reference like the default path."""
    g = golden["c2"]
    m = make("discogs-maest-10s-pw-129e", 62, fuse_ln=True)
    with torch.no_grad():
        lo, em = m(synth.wave_a(2, 160000).cuda())
        e6 = m(synth.wave_a(2, 160000).cuda(), transformer_block=6)[1]
    assert rel(lo, g["logits"]) < TOL_F16 and rel(em, g["emb"]) < TOL_F16
    m2 = make("discogs-maest-10s-pw-129e", 62)
    with torch.no_grad():
        lo2, _ = m2(synth.wave_a(2, 160000).cuda())
        e62 = m2(synth.wave_a(2, 160000).cuda(), transformer_block=6)[1]
    assert rel(lo2, lo) < 5e-4 and rel(e62, e6) < 5e-4
    with torch.no_grad():                      # no atomics anywhere: two runs are bit-identical
        assert torch.equal(m(synth.wave_a(2, 160000).cuda())[0], lo)


def test_attention_variants_agree(m10):
    x = synth.wave_a(2, 160000).cuda()
    default = m10.attn_variant
    try:
        with torch.no_grad():
            a = m10(x)[0]
            for v in (0, 1, 3):
                m10.attn_variant = v
                assert rel(a, m10(x)[0]) < 2e-4, v
    finally:
        m10.attn_variant = default


def test_config1_1d_clip_vs_reference(golden):
    g = golden["c1"]
    m = make("discogs-maest-10s-fs-129e", 62)
    with torch.no_grad():
        lo, em = m(synth.wave_a(1, 160000)[0].cuda())
    assert lo.shape == (1, 400) and em.shape == (1, 768)
    assert rel(lo, g["logits"]) < TOL_F16 and rel(em, g["emb"]) < TOL_F16


def test_input_rank_dispatch_vs_reference(m10, golden):
    g = golden["c2"]
    with torch.no_grad():
        lo, em = m10(synth.wave_a(1, 400000, seed=99)[0].cuda())          # 25 s -> 2 chunks
        assert lo.shape == (2, 400) and rel(lo, g["logits_25s_1d"]) < TOL_F16 and rel(em, g["emb_25s_1d"]) < TOL_F16
        lo, _ = m10(synth.wave_a(1, 48000, seed=98)[0].cuda())            # 3 s -> one short item
        assert lo.shape == (1, 400) and rel(lo, g["logits_3s_1d"]) < TOL_F16
        lo, _ = m10(synth.wave_b(160000)[None].cuda())
        assert rel(lo, g["logits_waveb"]) < TOL_F16
        gen = torch.Generator().manual_seed(5)
        m2 = torch.rand(96, 1300, generator=gen)
        m3 = torch.rand(2, 96, 625, generator=gen)
        assert rel(m10(m2.cuda(), melspectrogram_input=True)[0], g["logits_mel2d"]) < TOL_F16
        x3 = m3.clone().cuda()
        assert rel(m10(x3)[0], g["logits_mel3d"]) < TOL_F16
        assert x3.dim() == 4                                               # in-place unsqueeze_ quirk (models/maest.py:895)
        assert rel(m10(m3[:, None].cuda())[0], g["logits_mel4d"]) < TOL_F16
        # CPU input tensors are moved to the model's device
        assert rel(m10(m3[:, None].clone())[0], g["logits_mel4d"]) < TOL_F16
        # 2-D mel shorter than img_size[1] -> empty batch, no error
        lo, em = m10(torch.rand(96, 300).cuda(), melspectrogram_input=True)
        assert lo.shape == (0, 400) and em.shape == (0, 768)


def test_separated_heads_vs_reference(golden):
    g = golden["c2sep"]
    m = make("discogs-maest-10s-pw-129e", 62, distilled_type="separated")
    with torch.no_grad():
        lc, ld, ft = m(synth.wave_a(2, 160000).cuda())
    assert rel(lc, g["logits_cls"]) < TOL_F16 and rel(ld, g["logits_dist"]) < TOL_F16 and rel(ft, g["feats"]) < TOL_F16


def test_config3_30s_vs_reference(m30, golden):
    g = golden["c3"]
    x = synth.wave_a(2, 480000).cuda()
    with torch.no_grad():
        lo, em = m30(x)
    assert lo.shape == (2, 400)
    assert rel(lo, g["logits"]) < TOL_F16 and rel(em, g["emb"]) < TOL_F16
    assert np.abs(ops.logmel(x[:1]).cpu().numpy()[0][:, ::9] - g["mel0_sub"]).max() < 5e-5


def test_config5_predict_labels_vs_reference(golden):
    g = golden["c5"]
    m = make("discogs-maest-30s-pw-129e-519l", 187, n_classes=519)
    with torch.no_grad():
        act, labels = m.predict_labels(synth.wave_a(1, 95 * 16000, seed=77)[0].cuda())
        assert isinstance(act, np.ndarray) and act.dtype == np.float32 and act.shape == (519,) and len(labels) == 519
        assert np.abs(act - g["act_a"]).max() < 1e-3
        act_b, _ = m.predict_labels(synth.wave_b(95 * 16000).cuda())
        assert np.abs(act_b - g["act_b"]).max() < 1e-3
        act30, _ = m.predict_labels(synth.wave_a(1, 480000, seed=76)[0].cuda())
        assert np.abs(act30 - g["act_30s"]).max() < 1e-3
        e = m(synth.wave_a(1, 95 * 16000, seed=77)[0].cuda(), transformer_block=6)[1]
        assert e.shape == (3, 2304) and rel(e, g["emb6_a"]) < TOL_F16
        e = m(synth.wave_b(95 * 16000).cuda(), transformer_block=6)[1]
        assert rel(e, g["emb6_b"]) < TOL_F16


# ------------------------------------------------------------------------------------------ the reference's own API tests (tests/test_maest.py), ported
def test_reference_api_contract(m30):
    with pytest.raises(Exception):
        m30(np.random.rand(128, 128))
    with pytest.raises(Exception):
        m30(torch.empty([]))
    with pytest.raises(Exception, match="larger than the expected time encodings"):
        m30(torch.rand(2, 40 * 16000).float().cuda())
    with torch.no_grad():
        assert m30(torch.rand(10 * 16000).cuda())[0].shape == (1, 400)
        assert m30(torch.rand(2, 10 * 16000).cuda(), melspectrogram_input=False)[0].shape == (2, 400)
        assert m30(torch.rand(96, 1875).cuda(), melspectrogram_input=True)[0].shape == (1, 400)
        assert m30(torch.rand(96, 1875).cuda(), melspectrogram_input=True, transformer_block=6)[1].shape == (1, 2304)
        assert m30(torch.rand(2, 96, 1875).cuda(), melspectrogram_input=True, transformer_block=6)[1].shape == (2, 2304)
        assert m30(torch.rand(2, 1, 96, 1875).cuda(), melspectrogram_input=True, transformer_block=6)[1].shape == (2, 2304)
        # time-encoding limit: T <= 1885 (187 patches) fine, 1886 raises
        assert m30(torch.rand(1, 96, 1885).cuda())[0].shape == (1, 400)
        with pytest.raises(Exception):
            m30(torch.rand(1, 96, 1886).cuda())


# ------------------------------------------------------------------------------------------ full-size properties (BASELINE.json config 3: batch 64, 30 s)
def test_full_size_batch64_properties(m30):
    B = 64
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = torch.rand(B, 480000, generator=g, device="cuda") * 2 - 1
    with torch.no_grad():
        lo1, em1 = m30(x)
        lo2, em2 = m30(x)
        assert lo1.shape == (B, 400) and torch.isfinite(lo1).all()
        assert torch.equal(lo1, lo2) and torch.equal(em1, em2)                 # deterministic (no atomics / split-K)
        # clips are independent: a clip computed alone or at another batch position gives the same logits
        solo, _ = m30(x[5:6])
        assert rel(solo, lo1[5:6]) < 1e-5
        perm = torch.randperm(B, generator=torch.Generator().manual_seed(0)).cuda()
        lo_p, _ = m30(x[perm].contiguous())
        assert rel(lo_p, lo1[perm]) < 1e-5
        # a clip of the batch agrees with the CPU oracle (fp32) at full sequence length
        sd = synth.synth_state_dict(187)
        ref, _ = O.forward(x[:1].cpu(), sd, 1875, dtype=torch.float32)
        assert rel(lo1[:1], ref) < TOL_F16
        # predict_labels-style pooling is permutation invariant over the chunk axis
        assert rel(torch.sigmoid(lo_p).mean(0), torch.sigmoid(lo1).mean(0)) < 1e-5


@pytest.mark.gpu
def test_tensor_map_cache_serves_repeat_calls():
    """api.cu make_tmap: the second forward over the same buffers encodes no new TMA descriptors and returns identical bits."""
    import ctypes
    lib = _lib.init(0)
    model = get_maest(arch="discogs-maest-10s-pw-129e", pretrained=False)
    model.load_state_dict(synth.synth_state_dict(62, 400, seed=0), strict=False)
    model = model.cuda().eval()
    x = synth.wave_a(2, 160000).cuda()
    h, m = ctypes.c_uint64(), ctypes.c_uint64()
    with torch.no_grad():
        a, _ = model(x.clone())
        b, _ = model(x.clone())       # warm: workspace and 16-bit weight copies exist, torch's allocator reuses the blocks
        lib.maest_tmap_cache_stats(ctypes.byref(h), ctypes.byref(m))
        h0, m0 = h.value, m.value
        c, _ = model(x.clone())
        lib.maest_tmap_cache_stats(ctypes.byref(h), ctypes.byref(m))
    assert torch.equal(a, b) and torch.equal(b, c)
    assert h.value - h0 >= 100 and m.value - m0 <= 8, (h.value - h0, m.value - m0)


@pytest.mark.gpu
def test_wave_tokens_one_call_equals_k1_then_k2():
    """maest_wave_tokens_fwd (K1 + K2 behind one C-ABI entry) gives the bits of maest_logmel_fwd followed by
    maest_patch_tokens_fwd, with and without patchout."""
    model = get_maest(arch="discogs-maest-10s-pw-129e", pretrained=False, s_patchout_t=20, s_patchout_f=2)
    model.load_state_dict(synth.synth_state_dict(62, 400, seed=0), strict=False)
    model = model.cuda()
    x = synth.wave_a(3, 160000).cuda()
    args = (model._weight16("patch_embed.proj", model.patch_embed.proj.weight), model._f32(model.patch_embed.proj.bias),
            model._f32(model.freq_new_pos_embed).reshape(768, -1), model._f32(model.time_new_pos_embed).reshape(768, -1),
            model._f32(model.cls_token).reshape(-1), model._f32(model.dist_token).reshape(-1), model._f32(model.new_pos_embed).reshape(2, 768))
    mel = ops.logmel(x)
    assert torch.equal(ops.wave_tokens(x, *args), ops.patch_tokens(mel, *args))
    keep = ops.keep_ft_tensor(torch.tensor([0, 2, 3, 5, 6, 7, 8]), torch.arange(0, 62, 2), 9, 62, None, x.device)
    a = ops.wave_tokens(x, *args, keep_ft=keep)
    b = ops.patch_tokens(mel, *args, keep_ft=keep)
    with pytest.raises(Exception, match="larger than the expected time encodings"):      # models/maest.py:664-668
        ops.wave_tokens(x, *args, t_offset=3)
    assert a.shape == (3, 2 + 7 * 31, 768) and torch.equal(a, b)


@pytest.mark.gpu
def test_cuda_graph_forward_is_identical_and_tracks_weight_updates():
    """model.use_cuda_graphs (opt-in): replayed forwards return the eager bits for waveform and mel inputs and for a block
    embedding; an in-place weight update (new parameter version) drops the captured graphs."""
    model = get_maest(arch="discogs-maest-10s-pw-129e", pretrained=False)
    model.load_state_dict(synth.synth_state_dict(62, 400, seed=0), strict=False)
    model = model.cuda().eval()
    x2 = synth.wave_a(2, 160000).cuda()
    x1 = synth.wave_a(1, 21 * 16000)[0].cuda()              # 1-D, 21 s -> 2 chunks
    with torch.no_grad():
        ref2, ref1, ref6 = model(x2), model(x1), model(x2, transformer_block=6)
        model.use_cuda_graphs = True
        for _ in range(3):                                   # capture, then replays
            got2, got1, got6 = model(x2), model(x1), model(x2, transformer_block=6)
            assert torch.equal(got2[0], ref2[0]) and torch.equal(got2[1], ref2[1])
            assert torch.equal(got1[0], ref1[0]) and got6[0] is None and torch.equal(got6[1], ref6[1])
        y = model(x2 * 0.5)                                  # same shape, new data: replay reads the refreshed static input
        model.use_cuda_graphs = False
        assert torch.equal(y[0], model(x2 * 0.5)[0])
        model.use_cuda_graphs = True
        model.head[1].bias.add_(1.0)                         # version bump -> graphs dropped, new capture sees the new weights
        z = model(x2)
        assert float((z[0] - (ref2[0] + 1.0)).abs().max()) < 1e-5
    import copy
    assert copy.deepcopy(model) is not None                  # the graph cache is not copied (SWA deep-copies the net)
