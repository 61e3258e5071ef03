import sys, time, torch
sys.path.insert(0, '.')
from pose_splatter_b200 import _capi, batched, synth
dev = torch.device('cuda', 0)
d = synth.make_views('c2', 64, 6, seed=1)
p, vf, vm, Ks = (d[k].to(dev) for k in ('params', 'view_frame', 'viewmats', 'Ks'))
bg = torch.ones(3, device=dev)
W, H = d['width'], d['height']
w_rgb, w_a = synth.cotangents(len(vf), H, W, 7)
w_rgb, w_a = w_rgb.to(dev), w_a.to(dev)
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rgb, alpha, _, sv = batched.forward_raw('3d', p, vf, vm, Ks, bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD)
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    g = batched.backward_raw(sv, p, vf, vm, Ks, bg, w_rgb, w_a)
    t3 = time.perf_counter()
    sv.release()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f'fwd host {1e3*(t1-t0):.3f} ms (+sync {1e3*(t2-t1):.3f})  bwd host {1e3*(t3-t2):.3f} ms (+sync {1e3*(t4-t3):.3f})  total {1e3*(t4-t0):.3f}')
