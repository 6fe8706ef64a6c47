"""Live checks against the UNMODIFIED reference tree ($MAEST_REF, default /root/reference).  Container-only: the tree does not
exist on the GPU box, so every test here skips when it is absent (none is marked gpu)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not present")


def test_oracle_matches_the_live_reference_10s():
    """oracle/maest_oracle.py vs the reference itself, run now (models/maest.py:831-933), one 10 s clip, float32: the same bar as the
    committed fixtures (tests/test_oracle_golden.py)."""
    from maest_b200 import synth
    from oracle import maest_oracle as O
    ref = ref_loader.load_reference_maest()
    sd = synth.synth_state_dict(62, 400, seed=0)
    net = ref.get_maest(arch="discogs-maest-10s-pw-129e", pretrained=False)
    net.load_state_dict(sd, strict=False)
    net.eval()
    x = synth.wave_a(1, 160000)
    with torch.no_grad():
        lo_ref, emb_ref = net(x.clone())
        lo, emb = O.forward(x, sd, img_t=625, dtype=torch.float32)
    assert float((lo - lo_ref).norm() / lo_ref.norm()) < 2e-5
    assert float((emb - emb_ref).norm() / emb_ref.norm()) < 2e-5


def test_reference_specaugment_masking_is_a_no_op():
    """SURVEY.md section 8(f) row 1 names the SpecAugment masks of discogs/datamodule.py:140-152.  With the installed torchaudio (2.x)
    TimeMasking / FrequencyMasking are out-of-place and `masking_func` discards what `SpecMasking.compute` returns
    (helpers/spec_masking.py:27-33, datamodule.py:147-148), so the batch reaches the model UNMASKED: there is nothing for
    the ingest kernel to reproduce, and maest_b200/ingest.py deliberately has no masking stage."""
    sys.path.insert(0, ref_loader.REF_ROOT)
    try:
        from helpers.spec_masking import SpecMasking
    finally:
        sys.path.remove(ref_loader.REF_ROOT)
    torch.manual_seed(0)
    sm = SpecMasking(time_mask_param=8, freq_mask_param=5, p=0.2, iid_masks=True, time_masks=20, freq_masks=8)
    x = torch.randn(4, 1, 96, 1875)
    keep = x.clone()
    out = sm.compute(x)
    assert torch.equal(x, keep), "the reference's masking call modified its input in place: the ingest path must then mask too"
    assert not torch.equal(out, keep), "compute() itself does mask -- only its result is dropped by masking_func"
    # the exact statement sequence of datamodule.masking_func (discogs/datamodule.py:143-150)
    b = [x, ["f"] * 4, torch.zeros(4, 400)]
    xx = torch.as_tensor(b[0])
    sm.compute(xx)
    b[0] = xx
    assert torch.equal(b[0], keep)
