"""Small-batch latency of the public API (configs[0]/[4]-sized calls): one 10 s clip and one 30 s clip through model(x), host-timed
with a device sync per call (what a predict_labels user sees).  MAEST_TMAP_CACHE=0 disables the TMA-descriptor cache for an A/B."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from maest_b200 import get_maest, synth

out = {"tmap_cache": os.environ.get("MAEST_TMAP_CACHE", "1"), "cuda_graphs": os.environ.get("MAEST_CUDA_GRAPHS", "0")}
for arch, S, grid_t in (("discogs-maest-10s-fs-129e", 160000, 62), ("discogs-maest-30s-pw-129e", 480000, 187)):
    model = get_maest(arch=arch, pretrained=False)
    model.load_state_dict(synth.synth_state_dict(grid_t, 400, seed=0), strict=False)
    model = model.cuda().eval()
    model.use_cuda_graphs = os.environ.get("MAEST_CUDA_GRAPHS") == "1"
    for B in (1, 4):
        x = synth.wave_a(B, S).cuda()
        with torch.no_grad():
            for _ in range(5):
                model(x.clone())
            torch.cuda.synchronize()
            ts = []
            for _ in range(30):
                t0 = time.perf_counter()
                lo, _ = model(x.clone())
                lo[0, 0].item()
                ts.append(time.perf_counter() - t0)
        ts.sort()
        out[f"{arch} B={B}"] = dict(ms_median=round(1e3 * ts[len(ts) // 2], 3), ms_min=round(1e3 * ts[0], 3))
print("LATENCY " + json.dumps(out))
