"""Does splitting a step's frames over two CUDA streams pay?  python tools/overlap_probe.py [workload] [frames] [parts]
One call over all frames vs `parts` calls over contiguous frame ranges, each on its own stream (forward + backward).
3D workloads (frames shared by their cameras).  Measured at c2: 10.85 ms in one call, 11.18 ms on two streams, 12.04 on four."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from pose_splatter_b200 import _capi, batched, synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 256
parts = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
d = synth.make_views(wl, frames, 6, seed=3)
W, H, mode = d["width"], d["height"], d["mode"]
assert mode == "3d", "3D workloads only"
p, vf, vm, Ks = (d[k].to(dev) for k in ("params", "view_frame", "viewmats", "Ks"))
V = len(vf)
bg = torch.ones(3, device=dev)
w_rgb, w_a = synth.cotangents(V, H, W, seed=7)
w_rgb, w_a = w_rgb.to(dev), w_a.to(dev)


def whole():
    _, _, _, sv = batched.forward_raw(mode, p, vf, vm, Ks, bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD)
    g = batched.backward_raw(sv, p, vf, vm, Ks, bg, w_rgb, w_a)
    sv.release()
    return g


fpp = frames // parts
chunks = []
for c in range(parts):
    f0, f1 = c * fpp, (c + 1) * fpp if c < parts - 1 else frames
    sel = (vf >= f0) & (vf < f1)
    idx = sel.nonzero().flatten()
    chunks.append(dict(p=p[f0:f1].contiguous(), vf=(vf[idx] - f0).contiguous(), vm=vm[idx].contiguous() if vm is not None else None,
                       Ks=Ks[idx].contiguous() if Ks is not None else None, wr=w_rgb[idx].contiguous(), wa=w_a[idx].contiguous()))
streams = [torch.cuda.Stream(dev) for _ in range(parts)]


def split():
    main = torch.cuda.current_stream(dev)
    svs, gs = [], []
    for c, st in zip(chunks, streams):
        st.wait_stream(main)
        with torch.cuda.stream(st):
            _, _, _, sv = batched.forward_raw(mode, c["p"], c["vf"], c["vm"], c["Ks"], bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD)
            svs.append(sv)
    for c, st, sv in zip(chunks, streams, svs):
        with torch.cuda.stream(st):
            gs.append(batched.backward_raw(sv, c["p"], c["vf"], c["vm"], c["Ks"], bg, c["wr"], c["wa"]))
            sv.release()
    for st in streams:
        main.wait_stream(st)
    return gs


def timeit(fn, reps=6):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g0 = whole()
g1 = torch.cat(split(), 0)
print("gradients equal:", bool(torch.equal(g0, g1)), "max abs diff", float((g0 - g1).abs().max()))
print(f"{wl} {frames} frames: one call {timeit(whole):.3f} ms, {parts} calls on {parts} streams {timeit(split):.3f} ms")
