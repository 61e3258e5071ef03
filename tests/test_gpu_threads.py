"""One context per device is shared by every caller of the process (`_capi.context`): concurrent calls from several Python
threads on their own CUDA streams (ctypes drops the GIL inside the library) must not see each other -- every forward owns
its mailbox slot, the arena and the context's vectors are guarded, and a saved forward that is used or released on another
stream than the one that produced it is ordered behind the work already queued on it (round-1 ADVICE, item 1).

Compared with the same calls made one after the other on the default stream: images bit-identical, gradients to 1e-4 of
their column maximum (their atomic sums commute only up to rounding; measured differences are ~1e-6)."""
import threading

import numpy as np
import pytest
import torch

from helpers import column_rel_err

pytestmark = pytest.mark.gpu


def _jobs(dev):
    from pose_splatter_b200 import synth
    jobs = []
    # (workload, frames, cameras, N): a large 3D call (host wait on the list total, background fill on the context's side
    # stream), a small sync-free 3D call, and a 2D call
    for wl, frames, cams, n in (("c2", 2, 6, 16000), ("c2", 1, 2, 3000), ("c3", 1, 2, 2000)):
        d = synth.make_views(wl, n_frames=frames, n_cams=cams, seed=5 + len(jobs), n=n)
        V, H, W = len(d["view_frame"]), d["height"], d["width"]
        w_rgb, w_a = synth.cotangents(V, H, W, seed=11 + len(jobs))
        jobs.append(dict(mode=d["mode"], W=W, H=H, p=d["params"].to(dev), vf=d["view_frame"].to(dev),
                         vm=d["viewmats"].to(dev), Ks=d["Ks"].to(dev), bg=torch.tensor([0.2, 0.9, 0.5], device=dev),
                         w_rgb=w_rgb.to(dev), w_a=w_a.to(dev)))
    return jobs


def _run(job):
    from pose_splatter_b200 import batched
    return batched.render_views_vjp(job["mode"], job["p"], job["vf"], job["W"], job["H"], job["bg"], job["w_rgb"], job["w_a"],
                                    viewmats=job["vm"], Ks=job["Ks"])


def _same(got, want, what):
    rgb, alpha, g = (x.cpu().numpy() for x in got)
    assert np.array_equal(rgb, want[0]), f"{what}: rgb differs from the serial call"
    assert np.array_equal(alpha, want[1]), f"{what}: alpha differs from the serial call"
    err = column_rel_err(g, want[2]).max()
    assert err <= 1e-4, f"{what}: gradient differs from the serial call by {err:.2e}"


def test_concurrent_threads_share_one_context():
    dev = torch.device("cuda", 0)
    jobs = _jobs(dev)
    want = [tuple(x.cpu().numpy() for x in _run(j)) for j in jobs]
    torch.cuda.synchronize()
    errors = []
    start = threading.Barrier(2 * len(jobs))

    def worker(k, reps):
        try:
            stream = torch.cuda.Stream(device=dev)
            stream.wait_stream(torch.cuda.default_stream(dev))
            with torch.cuda.stream(stream):
                start.wait()
                for r in range(reps):
                    got = _run(jobs[k])
                    if r % 3 == 0 or r == reps - 1:
                        stream.synchronize()
                        _same(got, want[k], f"thread of job {k}, repetition {r}")
                stream.synchronize()
        except Exception as e:  # noqa: BLE001 -- reported by the main thread
            errors.append(f"job {k}: {type(e).__name__}: {e}")
            try:
                start.abort()
            except Exception:  # noqa: BLE001
                pass

    # two threads per job: the large call's threads interleave their host waits with the small calls' launches
    threads = [threading.Thread(target=worker, args=(k % len(jobs), 8 if k % len(jobs) == 0 else 24)) for k in range(2 * len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "a worker thread hangs"
    assert not errors, "; ".join(errors)


def test_saved_forward_used_and_released_on_other_streams():
    """Forward on stream A, backward on stream B, release on the default stream, and the next call (which takes the
    released blocks from the arena) right behind it: the library orders the streams itself."""
    from pose_splatter_b200 import _capi, batched
    dev = torch.device("cuda", 0)
    job = _jobs(dev)[0]
    want = tuple(x.cpu().numpy() for x in _run(job))
    torch.cuda.synchronize()
    a, b = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    for rep in range(4):
        with torch.cuda.stream(a):
            rgb, alpha, _, saved = batched.forward_raw(job["mode"], job["p"], job["vf"], job["vm"], job["Ks"], job["bg"],
                                                       job["W"], job["H"], _capi.FLAG_SAVE_FOR_BACKWARD)
        with torch.cuda.stream(b):  # no wait_stream(a): only the library's own ordering protects the saved buffers
            g = batched.backward_raw(saved, job["p"], job["vf"], job["vm"], job["Ks"], job["bg"], job["w_rgb"], job["w_a"])
        saved.release()             # default stream: the freed blocks go to the next call on it ...
        again = _run(job)           # ... which overwrites them while stream b may still be reading
        torch.cuda.synchronize()
        _same((rgb, alpha, g), want, f"cross-stream repetition {rep}")
        _same(again, want, f"call behind the release, repetition {rep}")
