"""maest_b200 — B200-native (sm_100a) hot path of palonso/MAEST behind the reference's own Python API.

    from maest_b200 import get_maest
    model = get_maest(arch="discogs-maest-30s-pw-129e", pretrained=False).cuda().eval()
    logits, embeddings = model(waveform)            # same call as the reference's `maest.get_maest`

`import maest_b200` does not touch CUDA; the first forward dlopens maest_b200/lib/libmaest_b200.so
(hand-written sm_100a kernels behind the C ABI in include/maest_b200.h) and fails loudly if it is missing.
"""
from .maest import MAEST, get_maest, maest_ing  # noqa: F401

__all__ = ["MAEST", "get_maest", "maest_ing"]
