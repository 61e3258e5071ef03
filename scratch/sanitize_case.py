import sys, torch
sys.path.insert(0, '.')
from pose_splatter_b200 import _capi, batched, synth
dev = torch.device('cuda', 0)
for wl, n in (('c2', 3000), ('c3', 1200)):
    d = synth.make_views(wl, 2, 3, seed=1, n=n)
    p, vf, vm, Ks = (d[k].to(dev) for k in ('params', 'view_frame', 'viewmats', 'Ks'))
    bg = torch.ones(3, device=dev)
    W, H = d['width'], d['height']
    w_rgb, w_a = synth.cotangents(len(vf), H, W, 7)
    rgb, alpha, cnt, sv = batched.forward_raw(d['mode'], p, vf, vm, Ks, bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD | _capi.FLAG_KEEP_BINNING, True)
    g = batched.backward_raw(sv, p, vf, vm, Ks, bg, w_rgb.to(dev), w_a.to(dev))
    torch.cuda.synchronize()
    print(wl, float(rgb.sum()), float(g.abs().sum()))
    sv.release()
