"""Device leg of formats.RenderSequenceWriter on a GPU box: c2 frames rendered to uint8 RGBA in batches, written through the
writer (pinned slabs, copy stream, writer thread) and compared with a plain .cpu() of the same renders.
    python tools/writer_probe.py [frames] [frames_per_batch]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pose_splatter_b200 import batched, formats, synth  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 240
B = int(sys.argv[2]) if len(sys.argv) > 2 else 48
dev = torch.device("cuda", 0)
d = synth.make_views("c2", n_frames=B, n_cams=6, seed=3, n=4000)
W, H, C = d["width"], d["height"], 6
p, vf, vm, Ks = d["params"].to(dev), d["view_frame"].to(dev), d["viewmats"].to(dev), d["Ks"].to(dev)
bg = torch.ones(3, device=dev)
fn = "/tmp/writer_probe.npy"


def batch(i):  # a different batch every time: the frames rolled by i
    return batched.render_views_rgba8("3d", torch.roll(p, i, 0), vf, W, H, bg, vm, Ks)


want = np.concatenate([batch(i).cpu().numpy().reshape(B, C, H, W, 4) for i in range(T // B)])
torch.cuda.synchronize()
t0 = time.perf_counter()
with formats.RenderSequenceWriter(fn, T, C, H, W, write_batch_frames=50) as out:
    for i in range(T // B):
        out.put(batch(i))
    t_put = time.perf_counter() - t0
t_all = time.perf_counter() - t0
got = np.asarray(formats.load_render_sequence(fn))
ok = bool(np.array_equal(got, want))
print(f"writer_probe: {T} frames x {C} cameras {W}x{H} ({got.nbytes / 1e6:.0f} MB): equal={ok}, puts returned after {t_put * 1e3:.1f} ms, "
      f"closed after {t_all * 1e3:.1f} ms ({T * C / t_all:.0f} views/s to disk, npy backend)")
assert ok
