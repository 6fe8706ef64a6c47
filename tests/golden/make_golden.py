"""Generate golden input/output fixtures by running the UNMODIFIED reference (palonso/MAEST) on CPU.

Run in the dev container only (needs /root/reference):   python tests/golden/make_golden.py
Writes tests/golden/*.npz (small: logits / embeddings / sub-sampled intermediates).  Weights and
inputs are NOT stored — they are regenerated bit-identically from seeds by `maest_b200.synth`.

The reference has no golden vectors of its own (tests/test_maest.py: shapes and exceptions only), so
these files are the pin for `oracle/maest_oracle.py` and, through it, for the CUDA path.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from maest_b200 import synth  # noqa: E402

ROW_PROBE = [0, 1, 2, 3, 61, 62, 63, 200, 311, 548, 559]   # token rows sampled per block (10 s model)


def build(ref, arch, grid_t, n_classes=400, seed=0, double=False, **kw):
    net = ref.get_maest(arch=arch, pretrained=False, n_classes=n_classes, **kw)
    sd = synth.synth_state_dict(grid_t, n_classes=net.num_classes, seed=seed)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("melspectrogram.") for k in missing), missing
    net.eval()
    if double:
        net = net.double()
    return net, sd


def np32(t):
    return t.detach().to(torch.float32).numpy()


def main():
    torch.set_num_threads(os.cpu_count())
    ref = ref_loader.load_reference_maest()
    out = {}

    # ---- config 1: 10s-fs, one 1-D clip -------------------------------------------------------
    net, _ = build(ref, "discogs-maest-10s-fs-129e", 62)
    x = synth.wave_a(1, 160000)[0]
    with torch.no_grad():
        lo, em = net(x.clone())
    net64, _ = build(ref, "discogs-maest-10s-fs-129e", 62, double=True)
    with torch.no_grad():
        lo64, em64 = net64(x.double())
    out["c1"] = dict(logits=np32(lo), emb=np32(em), logits64=lo64.numpy(), emb64=em64.numpy())

    # ---- config 2 (B=2 slice): 10s-pw, 2-D waves; intermediates --------------------------------
    net, sd = build(ref, "discogs-maest-10s-pw-129e", 62)
    x = synth.wave_a(2, 160000)
    probes = {}

    def hook_in(mod, args):
        probes["tokens"] = args[0].detach().clone()

    hs = [net.blocks[0].register_forward_pre_hook(hook_in)]
    for i, b in enumerate(net.blocks):
        hs.append(b.register_forward_hook(lambda m, a, o, i=i: probes.__setitem__(f"b{i}", o.detach().clone())))
    with torch.no_grad():
        mel = net.melspectrogram(x)
        lo, em = net(x.clone())
    for h in hs:
        h.remove()
    d = dict(logits=np32(lo), emb=np32(em), mel0=np32(mel[0]), mel1_sub=np32(mel[1, :, ::5]),
             tokens_probe=np32(probes["tokens"][:, ROW_PROBE, :]), row_probe=np.array(ROW_PROBE))
    for i in range(12):
        d[f"block{i}_probe"] = np32(probes[f"b{i}"][:, ROW_PROBE, :])
    with torch.no_grad():
        for k in (0, 6, 11):
            d[f"emb_block{k}"] = np32(net(x.clone(), transformer_block=k)[1])
        d["emb_block6_selfattn"] = np32(net(x.clone(), transformer_block=6, return_self_attention=True)[1])
        # 25 s of audio into the 10 s model -> 2 chunks (models/maest.py:868-875)
        lo25, em25 = net(synth.wave_a(1, 400000, seed=99)[0])
        d["logits_25s_1d"], d["emb_25s_1d"] = np32(lo25), np32(em25)
        # short 1-D wave (3 s) -> one short item (:876-877)
        lo3, em3 = net(synth.wave_a(1, 48000, seed=98)[0])
        d["logits_3s_1d"], d["emb_3s_1d"] = np32(lo3), np32(em3)
        # quiet tonal input
        lob, emb_ = net(synth.wave_b(160000)[None, :].clone())
        d["logits_waveb"], d["emb_waveb"] = np32(lob), np32(emb_)
        melb = net.melspectrogram(synth.wave_b(160000))
        d["mel_waveb"] = np32(melb)
        # mel inputs: 2-D (chunked), 3-D, 4-D
        g = torch.Generator().manual_seed(5)
        m2 = torch.rand(96, 1300, generator=g)
        d["logits_mel2d"] = np32(net(m2.clone(), melspectrogram_input=True)[0])
        m3 = torch.rand(2, 96, 625, generator=g)
        d["logits_mel3d"] = np32(net(m3.clone())[0])
        d["logits_mel4d"] = np32(net(m3.clone()[:, None])[0])
    out["c2"] = d

    # separated heads
    net_s, _ = build(ref, "discogs-maest-10s-pw-129e", 62, distilled_type="separated")
    with torch.no_grad():
        lc, ld, ft = net_s(x.clone())
    out["c2sep"] = dict(logits_cls=np32(lc), logits_dist=np32(ld), feats=np32(ft))

    # ---- config 3 (B=2 slice): 30s-pw, 2-D waves ------------------------------------------------
    net, _ = build(ref, "discogs-maest-30s-pw-129e", 187)
    x = synth.wave_a(2, 480000)
    with torch.no_grad():
        lo, em = net(x.clone())
        mel = net.melspectrogram(x)
    out["c3"] = dict(logits=np32(lo), emb=np32(em), mel0_sub=np32(mel[0, :, ::9]))

    # ---- config 5: 30s-pw-519l predict_labels / block-6 embeddings on 95 s 1-D --------------------
    net, _ = build(ref, "discogs-maest-30s-pw-129e-519l", 187)
    xa = synth.wave_a(1, 95 * 16000, seed=77)[0]
    xb = synth.wave_b(95 * 16000)
    with torch.no_grad():
        act_a, labels = net.predict_labels(xa.clone())
        act_b, _ = net.predict_labels(xb.clone())
        emb_a = net(xa.clone(), transformer_block=6)[1]
        emb_b = net(xb.clone(), transformer_block=6)[1]
        act30, _ = net.predict_labels(synth.wave_a(1, 480000, seed=76)[0])
    assert len(labels) == 519
    out["c5"] = dict(act_a=act_a, act_b=act_b, emb6_a=np32(emb_a), emb6_b=np32(emb_b), act_30s=act30)

    # ---- config 4: training step (loss + gradient probes), B=2 -----------------------------------
    import numpy.random as npr
    modm = ref_loader.load_reference_module()
    kwargs = dict(arch="passt_s_swa_p16_128_ap476", pretrained=False, n_classes=400, input_f=96,
                  input_t=1875, s_patchout_t=90)
    sd4 = synth.synth_state_dict(187, 400, seed=0)

    def _get():
        n = ref.get_maest(**kwargs)
        n.load_state_dict(sd4, strict=False)
        return n

    modm.get_maest = _get
    mod = modm.Module(do_swa=False, swa_epoch_start=50, swa_lrs=2e-5, swa_freq=5, mixup_alpha=0.3)
    mod.train()
    xb4, yb4 = synth.train_batch(2)
    torch.manual_seed(1)
    npr.seed(1)
    loss = mod.training_step((xb4.float(), ["a", "b"], yb4.float()), 0)
    loss.backward()
    # replay the host RNG draws in the reference's call order
    torch.manual_seed(1)
    npr.seed(1)
    rn = torch.randperm(2)
    lam = npr.beta(0.3, 0.3, 2).astype(np.float32)
    lam = np.concatenate([lam[:, None], 1 - lam[:, None]], 1).max(1)
    toff = int(torch.randint(1 + 187 - 186, (1,)).item())
    keep_t = torch.randperm(186)[: 186 - 90].sort().values
    g = {k: p.grad for k, p in mod.net.named_parameters()}
    d4 = dict(loss=np.float64(loss.item()), rn=rn.numpy(), lam=lam, toffset=np.int64(toff), keep_t=keep_t.numpy())
    for k in ["cls_token", "time_new_pos_embed", "freq_new_pos_embed", "patch_embed.proj.bias",
              "blocks.0.norm1.weight", "blocks.0.attn.qkv.bias", "blocks.5.mlp.fc1.bias",
              "blocks.11.attn.proj.bias", "norm.weight", "head.0.bias", "head.1.bias"]:
        d4["grad." + k] = np32(g[k])
    for k in ["patch_embed.proj.weight", "blocks.0.attn.qkv.weight", "blocks.5.mlp.fc1.weight",
              "blocks.11.mlp.fc2.weight", "head.1.weight"]:
        d4["grad." + k + ".sub"] = np32(g[k].reshape(g[k].shape[0], -1)[::37, ::29])
        d4["gnorm." + k] = np.float64(g[k].double().norm().item())
    assert g["head_dist.weight"] is None
    out["c4"] = d4

    # state-dict contract (names + shapes) of the reference models
    import json
    contract = {}
    for arch, n_cls in (("discogs-maest-30s-pw-129e", 400), ("discogs-maest-10s-fs-129e", 400), ("discogs-maest-5s-pw-129e", 400),
                        ("discogs-maest-20s-pw-129e", 400), ("discogs-maest-30s-pw-129e-519l", 400), ("passt_s_swa_p16_128_ap476", 400)):
        n = ref.get_maest(arch=arch, pretrained=False, n_classes=n_cls)
        contract[arch] = {k: list(v.shape) for k, v in n.state_dict().items()}
    with open(os.path.join(HERE, "state_dict_contract.json"), "w") as f:
        json.dump(contract, f)

    for name, d in out.items():
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **d)
        print(name, {k: getattr(v, "shape", None) for k, v in d.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
