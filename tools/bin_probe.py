"""Stage times of a few forward(+backward) passes at a reduced batch, for profiling runs (ncu) of single kernels.
    python tools/bin_probe.py [workload] [frames] [reps] [fwd|fwdbwd] [cameras]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pose_splatter_b200 import _capi, batched, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bwd = (sys.argv[4] if len(sys.argv) > 4 else "fwdbwd") == "fwdbwd"
cams = int(sys.argv[5]) if len(sys.argv) > 5 else 6
dev = torch.device("cuda", 0)
d = synth.make_views(wl, frames, cams, seed=3)
W, H, mode = d["width"], d["height"], d["mode"]
p, vf, vm, Ks = (d[k].to(dev) for k in ("params", "view_frame", "viewmats", "Ks"))
V = len(vf)
bg = torch.ones(3, device=dev)
w_rgb, w_a = synth.cotangents(V, H, W, seed=7)
w_rgb, w_a = w_rgb.to(dev), w_a.to(dev)
_capi.set_profiling(dev, True)
for it in range(reps):
    rgb, alpha, _, sv = batched.forward_raw(mode, p, vf, vm, Ks, bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD)
    if bwd:
        batched.backward_raw(sv, p, vf, vm, Ks, bg, w_rgb, w_a)
    info = sv.info()
    sv.release()
    torch.cuda.synchronize()
    st = _capi.stage_times(dev, reset=True)
    print(f"rep {it}: M={info.n_isect} lists={info.n_lists} " + " ".join(f"{k}={v[0]:.3f}" for k, v in st.items() if v[1]), flush=True)
