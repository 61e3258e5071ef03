#!/usr/bin/env python
"""bench.py -- fwd+bwd rendered views/s of the renderer hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c1|c5_3d|c5_2d] [--frames F]
    python bench.py --impl reference ...      # CPU arm: the oracle port on the box's host cores

One "step" = forward + backward of F frames x 6 cameras of synthetic Gaussians through the
C ABI (libpsplat.so).  `value` is measured with the inputs resident in HBM, on the device with
CUDA events, max over ranks; `e2e` goes through the public API (render_views + autograd) from
pinned HOST buffers with the host<->device copies inside the timed region.  N > 1: one process
per GPU (torchrun), whole frames sharded across ranks (weak scaling, no data-path collective);
--split-frames shards single views instead and all-reduces d_params over NCCL.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "fwd+bwd rendered views/s (6-cam, 576\u00d7512 2D / 288\u00d7256 3D GS) at 1/2/4/8 B200"
# SURVEY.md 8d-d4 pair model, split by what a pair actually costs: every evaluated pair pays the
# sigma / alpha evaluation (2 sub, 9 sigma, 2 exp scaling, o*e, min, 3 compares = 21 forward; + T recovery = 24 backward);
# only contributing pairs pay compositing (6 colour FMA-flops forward) or the gradient terms (36 backward)
FLOPS_EVAL = {"raster_fwd": 21.0, "raster_bwd": 24.0}
FLOPS_CONTRIB = {"raster_fwd": 6.0, "raster_bwd": 36.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", help="BASELINE.json config: c2 = 3D 288x256 (default), c3 = 2D 576x512")
    ap.add_argument("--frames", type=int, default=0, help="frames per step per GPU (x6 cameras = views); 0 = workload default")
    ap.add_argument("--split-frames", action="store_true", help="shard views (not frames): NCCL all-reduce of d_params")
    ap.add_argument("--fused-reduce", action="store_true",
                    help="with --split-frames: projection backward adds rows straight into the owner rank's d_params over NVLink")
    ap.add_argument("--forward-only", action="store_true",
                    help="inference: time the forward render alone (SURVEY 8d-d1: forward-only views/s for c1 and c4)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short c3 (2D) measurement that rides along with c2")
    ap.add_argument("--n", type=int, default=0, help="override Gaussians per frame (debug only; invalidates the metric)")
    return ap.parse_args()


FRAMES_DEFAULT = {"c1": 64, "c2": 256, "c3": 32, "c4": 256, "c5_3d": 16, "c5_2d": 4}
C4_SEQUENCE_FRAMES = 3600  # BASELINE.json configs[3]: full-sequence inference render, 3600 frames x 6 cameras


# ------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arm: oracle port on host cores (the reference has no native build; its 3D arithmetic is gsplat CUDA)
# ------------------------------------------------------------------------------------------
def cpu_views_per_second(workload, n_views, threads, n_override=0):
    """Times the CPU oracle (oracle/ps_oracle.c: the reference's algorithm restated in C) on `n_views` views
    of the workload, `threads` views in flight (ctypes releases the GIL). Returns (views/s, seconds)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as ora
    from pose_splatter_b200 import synth
    n_frames = max(1, (n_views + 5) // 6)
    d = synth.make_views(workload, n_frames, 6, seed=99, n=n_override or None)
    W, H, mode = d["width"], d["height"], d["mode"]
    w_rgb, w_a = synth.cotangents(n_views, H, W, seed=5)
    params = d["params"].numpy()
    vf = d["view_frame"].numpy()
    vms, Ks = d["viewmats"].numpy(), d["Ks"].numpy()
    bg = np.ones(3, np.float32)
    ora.lib()

    def one(v):
        ora.render(mode, params[int(vf[v])], W, H, bg, vms[v], Ks[v], w_rgb[v].numpy(), w_a[v].numpy())

    one(0)  # warm-up (page in the library, first-touch)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, range(n_views)))
    dt = time.perf_counter() - t0
    return n_views / dt, dt


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def reference_2d_cpu(wl, reps=3):
    """Baseline A of BASELINE.md section 3: the reference's OWN GaussianRenderer2D (baseline/_ref, byte-identical copy of
    src/gaussian_renderer.py made by __graft_entry__.build) on the box's host cores, built like the a6000_2d template
    (sigma_cutoff=3.0, kernel_size=5, batch_size=5), on camera 0 of the same synthetic workload.  Its autograd memory is
    ~N*H*W*40 B, so fwd+bwd is timed at N = 1024 (= min_n, src/model.py:32) and forward-only (no_grad) at N = 4096; the
    dense cost is exactly proportional to N (src/gaussian_renderer.py:379-425), which is what the labelled
    extrapolation to the workload's N uses."""
    import importlib.util
    import torch
    from pose_splatter_b200 import synth
    path = ROOT / "baseline" / "_ref" / "reference_src" / "gaussian_renderer.py"
    if not path.exists():
        return {"unavailable": "baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists"}
    spec = importlib.util.spec_from_file_location("reference_gaussian_renderer", str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = synth.WORKLOADS[wl]
    W, H, N = cfg["width"], cfg["height"], cfg["n"]
    d = synth.make_views(wl, 1, 6, seed=99)
    params = d["params"][0]
    r = mod.create_renderer("2d", W, H, device="cpu", sigma_cutoff=3.0, kernel_size=5, batch_size=5)
    r.set_background_color(torch.ones(3))
    w_rgb, w_a = synth.cotangents(1, H, W, seed=5)

    def fwd_bwd(n):
        p = params[:n].clone().requires_grad_(True)
        rgb, alpha = r.render(p, None, None)
        ((rgb * w_rgb[0]).sum() + (alpha * w_a[0]).sum()).backward()

    def fwd(n):
        with torch.no_grad():
            r.render(params[:n], None, None)

    def median_seconds(fn, n):
        fn(n)  # warm-up
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn(n)
            ts.append(time.perf_counter() - t0)
        return sorted(ts)[len(ts) // 2]

    n_bwd, n_fwd = min(1024, N), min(4096, N)
    t_bwd, t_fwd = median_seconds(fwd_bwd, n_bwd), median_seconds(fwd, n_fwd)
    return {"kind": "reference", "what": "reference GaussianRenderer2D (pure PyTorch, device='cpu'), one view per call",
            "cores": cores, "torch_threads": torch.get_num_threads(), "cpu": cpu_model_name(),
            "workload": f"{wl}: {W}x{H}, camera 0 of the synthetic workload, 1 warm-up + {reps} repetitions, median",
            "fwd_bwd": {"n": n_bwd, "seconds": t_bwd, "views_per_s": 1.0 / t_bwd,
                        "extrapolated_views_per_s_at_workload_n": (1.0 / t_bwd) * n_bwd / N, "workload_n": N,
                        "extrapolation": "linear in N (dense N*H*W cost); not a measurement"},
            "fwd_only": {"n": n_fwd, "seconds": t_fwd, "views_per_s": 1.0 / t_fwd,
                         "extrapolated_views_per_s_at_workload_n": (1.0 / t_fwd) * n_fwd / N, "workload_n": N,
                         "extrapolation": "linear in N (dense N*H*W cost); not a measurement" if n_fwd != N else "none: measured at the workload's N"}}


def reference_2d_on_gpu(wl, dev, reps=3):
    """Comparator of SURVEY 8d-d5: the reference's own GaussianRenderer2D run as it is with device='cuda' -- torch eager
    kernels on this B200 (baseline/_ref copy, unmodified; none of this repo's code on its path).  Forward only (no_grad) at
    the workload's N; forward + autograd backward at N = 1024 (its autograd memory is ~N*H*W*40 B)."""
    import importlib.util
    import torch
    from pose_splatter_b200 import synth
    path = ROOT / "baseline" / "_ref" / "reference_src" / "gaussian_renderer.py"
    if not path.exists():
        return {"unavailable": "baseline/_ref is missing: run __graft_entry__.build() where /root/reference exists"}
    spec = importlib.util.spec_from_file_location("reference_gaussian_renderer_gpu", str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = synth.WORKLOADS[wl]
    W, H, N = cfg["width"], cfg["height"], cfg["n"]
    params = synth.make_views(wl, 1, 6, seed=99)["params"][0].to(dev)
    r = mod.create_renderer("2d", W, H, device=str(dev), sigma_cutoff=3.0, kernel_size=5, batch_size=5)
    r.set_background_color(torch.ones(3, device=dev))
    w_rgb, w_a = synth.cotangents(1, H, W, seed=5)
    w_rgb, w_a = w_rgb[0].to(dev), w_a[0].to(dev)

    def fwd_bwd(n):
        p = params[:n].clone().requires_grad_(True)
        rgb, alpha = r.render(p, None, None)
        ((rgb * w_rgb).sum() + (alpha * w_a).sum()).backward()

    def fwd(n):
        with torch.no_grad():
            r.render(params[:n], None, None)

    def median_seconds(fn, n):
        fn(n)
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn(n)
            torch.cuda.synchronize(dev)
            ts.append(time.perf_counter() - t0)
        return sorted(ts)[len(ts) // 2]

    try:
        n_bwd = min(1024, N)
        t_fwd, t_bwd = median_seconds(fwd, N), median_seconds(fwd_bwd, n_bwd)
    except Exception as e:  # e.g. out of memory in the reference's dense temporaries
        return {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    return {"kind": "reference", "what": "reference GaussianRenderer2D (pure PyTorch, device='cuda': torch eager on this B200), one view per call",
            "workload": f"{wl}: {W}x{H}, camera 0 of the synthetic workload, 1 warm-up + {reps} repetitions, median, host wall clock around a device sync",
            "fwd_only": {"n": N, "seconds": t_fwd, "views_per_s": 1.0 / t_fwd},
            "fwd_bwd": {"n": n_bwd, "seconds": t_bwd, "views_per_s": 1.0 / t_bwd,
                        "extrapolated_views_per_s_at_workload_n": (1.0 / t_bwd) * n_bwd / N, "workload_n": N,
                        "extrapolation": "linear in N (dense N*H*W cost); not a measurement"}}


def gsplat_comparator():
    """Extra comparator of north_star: the reference's gsplat path on one B200.  gsplat is not vendored in the reference
    and not installed in this image (no network); the repo-root gsplat/ package is this repo's own shim, not gsplat."""
    import importlib.util
    saved_path, saved_mod = list(sys.path), {k: v for k, v in sys.modules.items() if k == "gsplat" or k.startswith("gsplat.")}
    try:
        sys.path[:] = [p for p in sys.path if p not in ("", ".", str(ROOT))]
        for k in saved_mod:
            del sys.modules[k]
        found = importlib.util.find_spec("gsplat") is not None
    except Exception:
        found = False
    finally:
        sys.path[:] = saved_path
        sys.modules.update(saved_mod)
    return "installed but not benchmarked" if found else "unavailable (gsplat is not installed in this image and cannot be fetched: no network)"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    wl = args.workload
    per_step = max(cores, 6) if wl in ("c1", "c2") else max(2, min(cores, 6))
    from pose_splatter_b200 import synth
    cfg = synth.WORKLOADS[wl]
    vals = []
    for _ in range(args.warmup):
        cpu_views_per_second(wl, min(per_step, cores), cores, args.n)
    t_total, v_total = 0.0, 0
    for _ in range(args.steps):
        vps, dt = cpu_views_per_second(wl, per_step, cores, args.n)
        vals.append(vps)
        t_total += dt
        v_total += per_step
    value = v_total / t_total
    sample = f"{per_step} views per step of workload {wl} (full size: N={args.n or cfg['n']}, {cfg['width']}x{cfg['height']}, fwd+bwd)"
    out = {"metric": METRIC, "value": value, "unit": "views/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
           "config": {"workload": workload_name(wl), "views_per_step": per_step, "mode": cfg["mode"]},
           "cpu_baseline": {"value": value, "unit": "views/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "CPU oracle port (oracle/ps_oracle.c) of the reference algorithm on all host cores; the reference "
                   "ships no native code and its 3D arithmetic (gsplat) is CUDA-only, so there is no reference CPU path for "
                   "the 3D configuration the metric is quoted on; the reference's own 2D class is timed beside it "
                   "(cpu_baseline_reference)",
           "gsplat_b200": gsplat_comparator()}
    if not args.no_cpu_baseline:
        out["cpu_baseline_reference"] = reference_2d_cpu("c3" if cfg["mode"] == "3d" else wl)
    print(json.dumps(out), flush=True)


def workload_name(wl):
    from pose_splatter_b200 import synth
    c = synth.WORKLOADS[wl]
    idx = {"c1": 0, "c2": 1, "c3": 2, "c4": 3, "c5_3d": 4, "c5_2d": 4}[wl]
    what = "GS full-sequence inference render (forward only, uint8 RGBA out), 3600 frames x" if wl == "c4" else "GS fwd+bwd,"
    return f"BASELINE.json configs[{idx}] ({wl}): {c['mode'].upper()} {what} 6 cameras at {c['width']}x{c['height']}, N={c['n']} synthetic Gaussians per frame"


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def pin_rank_to_cores(local, local_world):
    """One process per GPU: give every rank a disjoint set of host cores on its GPU's NUMA node (its pinned staging
    buffers are allocated afterwards, first-touch on that node).  Eight un-pinned ranks sharing every core was the e2e
    limiter of round 1.  Returns what was done, for the JSON line."""
    import torch
    try:
        avail = sorted(os.sched_getaffinity(0))
    except Exception:
        return {"pinned": False, "why": "sched_getaffinity unavailable"}

    def node_of(dev_index):
        try:
            pr = torch.cuda.get_device_properties(dev_index)
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            return int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        except Exception:
            return -1

    def cpus_of(node):
        out = []
        try:
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                a, _, b = part.partition("-")
                out += list(range(int(a), int(b or a) + 1))
        except Exception:
            pass
        return [c for c in out if c in avail]

    nodes = [node_of(i) for i in range(local_world)]
    mine = nodes[local] if local < len(nodes) else -1
    pool = cpus_of(mine) if mine >= 0 else []
    peers = [i for i in range(local_world) if nodes[i] == mine] if pool else list(range(local_world))
    if not pool:
        pool = avail
    k = peers.index(local) if local in peers else 0
    per = max(1, len(pool) // max(1, len(peers)))
    cores = pool[k * per:(k + 1) * per] or pool
    try:
        os.sched_setaffinity(0, cores)
        torch.set_num_threads(max(1, min(len(cores), 8)))
    except Exception as e:
        return {"pinned": False, "why": str(e)}
    return {"pinned": True, "numa_node": mine, "cores": f"{cores[0]}-{cores[-1]}" if cores else "", "n_cores": len(cores),
            "host_cores_total": len(avail)}


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: pose_splatter_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pin = pin_rank_to_cores(local, int(os.environ.get("LOCAL_WORLD_SIZE", str(world))))
    args._pin = pin
    # stdout carries exactly ONE JSON line: anything a library prints on fd 1 meanwhile (NCCL's version banner at
    # communicator creation) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
        out = run_measurements(args, world, rank, local, dev)
        if world > 1:
            dist.destroy_process_group()
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if rank == 0:
        print(json.dumps(out), flush=True)


def split_frames_leg(args, world, rank, dev, steps=5):
    """BASELINE.json configs[4] layout at c2 size, part of the default N > 1 line: one frame's six cameras land on
    DIFFERENT ranks (view v -> rank v % world), so the per-Gaussian gradients of a frame must be summed across GPUs.
    Two ways, same inputs: (a) local backward + NCCL all-reduce of d_params over NVLink, (b) the fused exchange (K8'):
    the projection-backward kernel pushes finished rows into the owner rank's staging buffer over peer memory
    (ps_backward_peer) and the owner sums its slots.  The two gradients are compared on every rank."""
    import torch
    import torch.distributed as dist
    from pose_splatter_b200 import _capi, batched, synth
    from pose_splatter_b200 import dist as psd
    F, C = FRAMES_DEFAULT["c2"], 6
    d = synth.make_views("c2", F, C, seed=4242, n=args.n or None)  # the same frames on every rank
    W, H = d["width"], d["height"]
    mine = torch.tensor(psd.shard_views(F, C, rank, world, "view"), dtype=torch.long)
    p = d["params"].to(dev)
    vf, vm, Ks = d["view_frame"][mine].to(dev), d["viewmats"][mine].to(dev), d["Ks"][mine].to(dev)
    w_rgb, w_a = synth.cotangents(F * C, H, W, seed=9)
    w_rgb, w_a = w_rgb[mine].to(dev).contiguous(), w_a[mine].to(dev).contiguous()
    bg = torch.ones(3, device=dev)
    peer = psd.PeerGradBuffers(tuple(p.shape), dev)

    def step_nccl():
        _, _, _, sv = batched.forward_raw("3d", p, vf, vm, Ks, bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD)
        g = batched.backward_raw(sv, p, vf, vm, Ks, bg, w_rgb, w_a)
        sv.release()
        psd.reduce_frame_grads(g)
        return g

    def step_fused():
        _, _, _, sv = batched.forward_raw("3d", p, vf, vm, Ks, bg, W, H, _capi.FLAG_SAVE_FOR_BACKWARD)
        peer.begin()
        batched.backward_peer_raw(sv, p, vm, Ks, bg, w_rgb, w_a, peer.rank_ptrs, peer.rank, peer.world)
        g = peer.end()
        sv.release()
        return g

    res = {}
    grads = {}
    for name, fn in (("nccl_allreduce", step_nccl), ("fused_peer_push", step_fused)):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            g = fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        ms = psd.max_over_ranks(e0.elapsed_time(e1), dev) / steps
        grads[name] = g.clone()
        res[name] = {"value": F * C / (ms * 1e-3), "unit": "views/s", "ms_per_step": ms}
    want = grads["nccl_allreduce"][peer.owned]              # complete gradient of the frames this rank owns
    got = grads["fused_peer_push"]
    scale = want.abs().amax(dim=(0, 1)).clamp_min(1e-20)
    rel = float(((got - want).abs().amax(dim=(0, 1)) / scale).max()) if len(peer.owned) else 0.0
    rel = psd.max_over_ranks(rel, dev)
    res["max_rel_difference_of_the_two_gradients"] = rel
    res["gradients_agree"] = bool(rel < 1e-4)
    res["config"] = {"workload": workload_name("c2"), "frames_per_step_total": F, "views_per_step_total": F * C,
                     "sharding": f"view v -> rank v % {world}: every frame's cameras are split across ranks",
                     "exchange_bytes_per_step_per_rank": int(p.numel() * 4), "steps": steps}
    return res


def run_measurements(args, world, rank, local, dev):
    out = measure(args, args.workload, args.steps, world, rank, local, dev, primary=True)
    if world > 1 and args.workload == "c2" and not args.split_frames and not args.forward_only:
        # driver-visible evidence for the cross-GPU gradient exchange: part of every default N > 1 line
        leg = split_frames_leg(args, world, rank, dev)
        if rank == 0:
            out.setdefault("also", {})["split_frames"] = leg
    if args.workload == "c2" and not args.no_secondary and not args.n and not args.split_frames and not args.forward_only:
        # the metric names two configurations; the 2D one (BASELINE.json configs[2]) rides along, briefly
        other = measure(args, "c3", max(3, min(6, args.steps)), world, rank, local, dev, primary=False)
        if rank == 0:
            out.setdefault("also", {})["c3"] = {k: other[k] for k in ("value", "unit", "ms_per_step", "steps", "config", "e2e", "roofline",
                                                                      "stage_ms_per_step", "pairs", "per_kernel")}
    return out


def per_kernel_traffic(wl, mode, V, N, M, n_lists, stats_all, stage_ms):
    """Algorithmic bytes per launch of every HBM-relevant kernel (formulas of DESIGN.md section 7) beside the DRAM bytes of
    the committed ncu capture (profiles/ncu_traffic.json) and the live CUDA-event time of its stage."""
    VN = V * N
    P = 14 if mode == "3d" else 9
    f, b = stats_all["fwd"], stats_all["bwd"]
    rows_in = VN * 4 * P // (6 if mode == "3d" else 1)   # 3D: six cameras share a frame's rows
    alg = {
        "project": rows_in + VN * (64 + 8 + 4 + (4 if mode == "3d" else 0)),
        "partition": VN * (4 + 8 + 32 + (4 if mode == "3d" else 0)) + M * 4,
        "sort_lists": M * (8 if mode == "3d" else 4) + M * 4,
        "block_lists": M * 36 + f["entries_staged"] * 4,
        # forward: id + record per staged block-list entry, (id, pixel mask) written per contributor-list entry (= what the
        # backward stages); backward: id + mask + record per contributor-list entry
        "raster_fwd": f["entries_staged"] * 52 + b["entries_staged"] * 8 + n_lists * 256 * 28,
        "raster_bwd": b["entries_staged"] * 56 + n_lists * 256 * 24 + b["entries_walked"] * 36,
        "project_bwd": VN * (4 + 36 + 64) + 2 * rows_in,
    }
    stage_of = {"project": "project", "partition": "partition", "sort_lists": "sort", "block_lists": "blocks", "raster_fwd": "raster_fwd",
                "raster_bwd": "raster_bwd", "project_bwd": "project_bwd"}
    out = {}
    for k, a in alg.items():
        dram = ncu_traffic(wl, k)
        ms = stage_ms[stage_of[k]][0]
        out[k] = {"alg_bytes": int(a), "dram_bytes": dram, "ratio": (dram / a) if (dram and a) else None,
                  "stage_ms": ms, "alg_gbs": a / (ms * 1e-3) / 1e9 if ms else None}
    return out


def ncu_traffic(wl, kernel, field="dram_bytes_per_launch"):
    """DRAM bytes per launch (or another field) of `kernel` from the committed ncu --set full capture of this workload, or None."""
    try:
        return json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())[wl][kernel][field]
    except Exception:
        return None


def measure(args, wl, steps, world, rank, local, dev, primary):
    import torch
    import torch.distributed as dist
    from pose_splatter_b200 import _capi, batched, synth
    from pose_splatter_b200 import dist as psd

    cfg = synth.WORKLOADS[wl]
    mode, W, H = cfg["mode"], cfg["width"], cfg["height"]
    F = (args.frames if primary else 0) or FRAMES_DEFAULT[wl]
    n_cams = 6
    n_sets = 4  # rotate distinct input batches so no step finds its inputs in L2

    # --- synthetic inputs: rank r renders frames {r, r+world, ...} of a global sequence (weak scaling)
    sets = []
    for k in range(n_sets):
        d = synth.make_views(wl, F, n_cams, seed=1000 * rank + k, n=args.n or None)
        if args.split_frames and world > 1:
            # every rank holds all frames' parameters but renders only its views (v % world == rank)
            d = synth.make_views(wl, F, n_cams, seed=k, n=args.n or None)
            mine = torch.tensor(psd.shard_views(F, n_cams, rank, world, "view"), dtype=torch.long)
            if mode == "3d":
                d["view_frame"], d["viewmats"], d["Ks"] = d["view_frame"][mine], d["viewmats"][mine], d["Ks"][mine]
            else:
                d["params"], d["view_frame"] = d["params"][mine], torch.arange(len(mine)).int()
                d["viewmats"], d["Ks"] = d["viewmats"][mine], d["Ks"][mine]
        sets.append(d)
    V = int(sets[0]["view_frame"].shape[0])
    host = [dict(params=s["params"].pin_memory(), view_frame=s["view_frame"].pin_memory(),
                 viewmats=s["viewmats"].pin_memory(), Ks=s["Ks"].pin_memory()) for s in sets]
    devs = [{k: v.to(dev) for k, v in h.items()} for h in host]
    bg = torch.ones(3, device=dev)
    w_rgb, w_a = synth.cotangents(V, H, W, seed=7)
    w_rgb, w_a = w_rgb.to(dev), w_a.to(dev)
    need_reduce = args.split_frames and world > 1 and mode == "3d"
    fused = need_reduce and args.fused_reduce
    peer = psd.PeerGradBuffers(tuple(sets[0]["params"].shape), dev) if fused else None
    input_mb = sum(sum(t.numel() * t.element_size() for t in h.values()) for h in host) / 2**20

    fwd_only = bool(args.forward_only)
    img_host = None

    def step_resident(k):
        s = devs[k % n_sets]
        if fwd_only:  # inference writes uint8 RGBA like scripts/utils/evaluate_model.py:101-113
            return batched.render_views_rgba8(mode, s["params"], s["view_frame"], W, H, bg, s["viewmats"], s["Ks"])
        rgb, alpha, _, saved = batched.forward_raw(mode, s["params"], s["view_frame"], s["viewmats"], s["Ks"], bg, W, H,
                                                   _capi.FLAG_SAVE_FOR_BACKWARD)
        if fused:
            peer.begin()
            batched.backward_peer_raw(saved, s["params"], s["viewmats"], s["Ks"], bg, w_rgb, w_a, peer.rank_ptrs, peer.rank, peer.world)
            g_owned = peer.end()
            saved.release()
            return g_owned
        d_params = batched.backward_raw(saved, s["params"], s["view_frame"], s["viewmats"], s["Ks"], bg, w_rgb, w_a)
        saved.release()
        if need_reduce:
            psd.reduce_frame_grads(d_params)
        return d_params

    # --- end to end through the public API (render_views_vjp) from pinned host buffers.  Copies run on two
    # side streams so that step k+1's host->device copy and step k-1's device->host copy overlap step k's kernels
    # (every copy of every timed step is inside the timed region; the region ends when all three streams are idle).
    out_host = [torch.empty_like(host[0]["params"]).pin_memory() for _ in range(2)]
    loss_host = torch.empty(2).pin_memory()
    gnorm_host = [torch.empty(host[0]["params"].shape[0]).pin_memory() for _ in range(2)]
    e2e_cfg = {"grad_to_host": False}
    h2d_stream, d2h_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    staged = {}

    # three persistent device staging sets, reused round-robin: no allocator traffic inside the timed loop (a
    # cudaMalloc issued by the caching allocator for a side stream synchronises the device and stalls the pipeline:
    # measured as a step time that doubled on some runs)
    n_stage = 3
    stage_dev = [{name: torch.empty_like(v, device=dev) for name, v in host[0].items()} for _ in range(n_stage)]
    stage_free = [None] * n_stage  # event recorded on the main stream after the step that consumed the set

    def stage_inputs(k):
        j = k % n_stage
        with torch.cuda.stream(h2d_stream):
            if stage_free[j] is not None:
                h2d_stream.wait_event(stage_free[j])
            for name, v in host[k % n_sets].items():
                stage_dev[j][name].copy_(v, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(h2d_stream)
        staged[k] = (stage_dev[j], ev, j)

    def step_e2e(k, last=False):
        main = torch.cuda.current_stream(dev)
        if k not in staged:
            stage_inputs(k)
        t, ev, slot_j = staged.pop(k)
        if not last:
            stage_inputs(k + 1)
        main.wait_event(ev)
        if fwd_only:  # inference: rendered images go back to the host
            nonlocal img_host
            img = batched.render_views_rgba8(mode, t["params"], t["view_frame"], W, H, bg, t["viewmats"], t["Ks"])
            if img_host is None:
                img_host = [torch.empty_like(img, device="cpu").pin_memory() for _ in range(2)]
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                img_host[k % 2].copy_(img, non_blocking=True)
            img.record_stream(d2h_stream)
            stage_free[slot_j] = done
            return
        # public API: forward + vector-Jacobian product with the metric's fixed cotangents (no autograd graph: the
        # cotangent of L = sum(w_rgb * rgb) + sum(w_a * alpha) is w itself), loss value from the rendered images
        rgb, alpha, g = batched.render_views_vjp(mode, t["params"], t["view_frame"], W, H, bg, w_rgb, w_a, t["viewmats"], t["Ks"])
        loss = torch.dot(rgb.reshape(-1), w_rgb.reshape(-1)) + torch.dot(alpha.reshape(-1), w_a.reshape(-1))
        if need_reduce:
            psd.reduce_frame_grads(g)
        lossd = loss.detach().reshape(1)
        # the step's result goes back to the host: the loss and the per-frame gradient norms (what a training loop logs;
        # the gradient itself stays on the device for the optimiser, as in the reference's training step).  With
        # grad_to_host the whole d_params [F,N,P] is copied back as well (round-1 definition, kept as a secondary number).
        gnorm = torch.linalg.vector_norm(g.reshape(g.shape[0], -1), dim=1)
        done = torch.cuda.Event()
        done.record(main)
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(done)
            if e2e_cfg["grad_to_host"]:
                out_host[k % 2].copy_(g, non_blocking=True)
            gnorm_host[k % 2].copy_(gnorm, non_blocking=True)
            loss_host[k % 2:k % 2 + 1].copy_(lossd, non_blocking=True)
        g.record_stream(d2h_stream)
        gnorm.record_stream(d2h_stream)
        lossd.record_stream(d2h_stream)
        stage_free[slot_j] = done

    def drain_e2e():
        torch.cuda.current_stream(dev).wait_stream(d2h_stream)
        torch.cuda.current_stream(dev).wait_stream(h2d_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False, e2e=False):
        barrier()
        if profile:
            _capi.set_profiling(dev, True)
            _capi.stage_times(dev, reset=True)
        l0 = _capi.launch_count(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(steps):
            if e2e:
                fn(k, last=(k == steps - 1))
            else:
                fn(k)
        if e2e:
            drain_e2e()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        stages = None
        if profile:
            stages = _capi.stage_times(dev, reset=True)
            _capi.set_profiling(dev, False)
        return psd.max_over_ranks(ms, dev), _capi.launch_count(dev) - l0, stages

    for k in range(max(3, args.warmup)):
        step_resident(k)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total, launches, stages = timed(step_resident, steps, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    for k in range(max(3, args.warmup)):
        step_e2e(k, last=True)
    drain_e2e()
    ms_e2e, _, _ = timed(step_e2e, steps, e2e=True)
    ms_e2e_grad = None
    if not fwd_only:  # secondary (round-1 definition): the whole gradient d_params [F,N,P] is copied back as well
        e2e_cfg["grad_to_host"] = True
        for k in range(max(3, args.warmup)):  # the gradient tensors held by the copy stream: the allocator needs a few steps to settle
            step_e2e(k, last=True)
        drain_e2e()
        ms_e2e_grad, _, _ = timed(step_e2e, steps, e2e=True)
        ms_e2e_grad /= steps
        e2e_cfg["grad_to_host"] = False
    # copies alone (no kernels), all ranks at once: what the host side can deliver per step
    def copies_only(k, last=False):
        with torch.cuda.stream(h2d_stream):
            for name, v in host[k % n_sets].items():
                stage_dev[k % n_stage][name].copy_(v, non_blocking=True)
    for k in range(2):
        copies_only(k)
    drain_e2e()
    ms_copy, _, _ = timed(copies_only, steps, e2e=True)
    ms_copy_out = None
    if not fwd_only:  # device->host alone: the gradient-sized buffer into pinned host memory
        g_dev = torch.zeros_like(devs[0]["params"])
        def copy_out_only(k, last=False):
            with torch.cuda.stream(d2h_stream):
                out_host[k % 2].copy_(g_dev, non_blocking=True)
        copy_out_only(0)
        drain_e2e()
        ms_copy_out, _, _ = timed(copy_out_only, steps, e2e=True)

    # pair counts of one step (untimed): the SAME forward / backward kernels with their counters compiled in
    s0 = devs[0]
    _capi.raster_stats(dev, reset=True)
    _, _, _, sv = batched.forward_raw(mode, s0["params"], s0["view_frame"], s0["viewmats"], s0["Ks"], bg, W, H,
                                      _capi.FLAG_SAVE_FOR_BACKWARD | _capi.FLAG_RASTER_STATS)
    if not fwd_only:
        batched.backward_raw(sv, s0["params"], s0["view_frame"], s0["viewmats"], s0["Ks"], bg, w_rgb, w_a)
    stats_all = _capi.raster_stats(dev, reset=True)
    info = sv.info()
    M = int(info.n_isect)
    n_lists = int(info.n_lists)
    sv.release()
    fp32_peak = _capi.fp32_peak_tflops(dev)

    if rank != 0:
        return None
    h2d_bytes = int(sum(t.numel() * t.element_size() for t in host[0].values()))
    views_total = V * world * steps
    value = views_total / (ms_total * 1e-3)
    e2e_value = views_total / (ms_e2e * 1e-3)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    stage_ms = {k: (v[0] / max(1, v[1]), v[1]) for k, v in stages.items()}  # average per launch
    per_step = {k: v[0] / steps for k, v in stages.items()}
    dom = "raster_fwd" if fwd_only else max(("raster_fwd", "raster_bwd"), key=lambda k: per_step[k])
    dom_ms = stage_ms[dom][0]
    stats = stats_all["bwd" if dom == "raster_bwd" else "fwd"]  # counters of the dominant kernel itself
    # FP32 work of the pair model: every evaluated pair pays the alpha evaluation, only the
    # contributing ones pay the compositing / gradient terms
    flops = stats["pairs_evaluated"] * FLOPS_EVAL[dom] + stats["pairs_contributing"] * FLOPS_CONTRIB[dom]
    achieved_tf = flops / (dom_ms * 1e-3) / 1e12
    kname = {"raster_fwd": "raster_fwd6_kernel" if mode == "3d" else "raster_fwd_kernel", "raster_bwd": "raster_bwd3_kernel"}[dom]
    roofline = {"bound": "fp32", "kernel": dom, "cuda_kernel": kname, "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": achieved_tf / fp32_peak if fp32_peak else None, "traffic": ncu_traffic(wl, dom),
                "peak_source": "FFMA micro-benchmark run in this process (ps_fp32_peak_probe), 2 flops per FMA; "
                               "MEASURED_PEAKS.json has no FP32 figure; no tensor cores on this path",
                "work": f"counters of {kname} itself (same kernel, STATS template path): {stats['pairs_evaluated']} evaluated pairs x "
                        f"{FLOPS_EVAL[dom]:.0f} + {stats['pairs_contributing']} contributing pairs x {FLOPS_CONTRIB[dom]:.0f} FP32 ops per launch",
                "avg_launch_ms": dom_ms, "share_of_step": per_step[dom] / (ms_total / steps),
                # what actually bounds the kernel (committed ncu capture of the same kernel at this batch size): the share of
                # cycles in which a scheduler issues an instruction, and the active lanes per issued instruction
                "issue_bound": {"issue_active_frac": (ncu_traffic(wl, dom, "issue_active_pct") or 0.0) / 100.0 or None,
                                "active_lanes_per_instruction": ncu_traffic(wl, dom, "threads_per_inst"),
                                "warp_instructions_per_launch": ncu_traffic(wl, dom, "inst_executed"),
                                "note": "the rasterizers are bound by instruction issue on divergent lanes (per-pixel / per-entry walks over "
                                        "32-entry chunks), not by the FMA pipe or HBM: `frac` counts useful pair flops only"}}
    if not fwd_only:
        # the other rasterizer on the same terms (round 1's dominant kernel was the backward one: its fraction is tracked
        # across rounds whichever of the two is slower now)
        oth = "raster_bwd" if dom == "raster_fwd" else "raster_fwd"
        ost = stats_all["bwd" if oth == "raster_bwd" else "fwd"]
        oth_ms = stage_ms[oth][0]
        oth_tf = (ost["pairs_evaluated"] * FLOPS_EVAL[oth] + ost["pairs_contributing"] * FLOPS_CONTRIB[oth]) / (oth_ms * 1e-3) / 1e12
        roofline["other_rasterizer"] = {"kernel": oth, "achieved": oth_tf, "unit": "TFLOP/s", "frac": oth_tf / fp32_peak if fp32_peak else None,
                                        "avg_launch_ms": oth_ms, "traffic": ncu_traffic(wl, oth),
                                        "work": f"{ost['pairs_evaluated']} evaluated pairs x {FLOPS_EVAL[oth]:.0f} + "
                                                f"{ost['pairs_contributing']} contributing pairs x {FLOPS_CONTRIB[oth]:.0f} FP32 ops per launch"}
    # the same kernel against the HBM roofline: per staged entry 4 B id + 48 B record (+ 4 B pixel mask in the backward, which
    # stages the forward's contributor list; the forward writes 8 B per entry of that list), per pixel OF A NON-EMPTY TILE
    # 24 B (saved state + cotangents; forward: 20 B written + 8 B saved), 36 B of atomics per contributing entry (backward)
    dom_bytes = stats["entries_staged"] * (56 if dom == "raster_bwd" else 52) + n_lists * 256 * (24 if dom == "raster_bwd" else 28) + \
        (stats["entries_walked"] * 36 if dom == "raster_bwd" else stats_all["bwd"]["entries_staged"] * 8)
    roofline["as_hbm"] = {"bound": "hbm", "achieved": dom_bytes / (dom_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                          "frac": dom_bytes / (dom_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src, "algorithmic_bytes": dom_bytes,
                          "note": "the rasterizers are instruction-issue bound, not HBM bound: this is how far below the HBM roofline they sit"}
    VN = V * (args.n or cfg["n"])
    # rank: depth word + listed flag in, order + rank out; partition: rect + count + rank + 32 B of record in, 4 B slot word
    # out per entry; sort + split: slot word in, order gather, 4 B per block-list entry out; scan: one int in/out per
    # (view, tile)
    sort_bytes = ((VN * 16 if mode == "3d" else 0) + VN * (16 + 32) + M * 4 + M * (8 if mode == "3d" else 4)
                  + stats_all["fwd"]["entries_staged"] * 4 + 8 * V * info.tiles_x * info.tiles_y)
    sort_ms = stage_ms["rank"][0] + stage_ms["scan"][0] + stage_ms["partition"][0] + stage_ms["sort"][0] + stage_ms["blocks"][0]
    proj_bytes = VN * ((56 if mode == "3d" else 36) + 48 + 8 + 4)
    binning = {"kernels": "depth_rank+scan+partition+sort_split", "achieved": sort_bytes / (sort_ms * 1e-3) / 1e9 if sort_ms else None,
               "unit": "GB/s", "frac": (sort_bytes / (sort_ms * 1e-3) / 1e9) / hbm_peak if sort_ms else None, "ms_per_step": sort_ms,
               "work": f"M={M} list entries in {n_lists} non-empty (view,tile) lists, {VN} (view,Gaussian) records",
               "project": {"achieved": proj_bytes / (stage_ms['project'][0] * 1e-3) / 1e9 if stage_ms['project'][0] else None,
                           "unit": "GB/s", "avg_launch_ms": stage_ms["project"][0]}}
    # the HBM-bound kernel of the path: block_lists (default binning mode "gather").  Algorithmic bytes per tile-list entry:
    # 4 (list id) + 32 (the two cull words of the splat record) + 4 per block-list entry written
    blk_ms = stage_ms["blocks"][0]
    blk_entries = stats_all["fwd"]["entries_staged"]
    blk_bytes = M * 36 + blk_entries * 4
    blk_traffic = ncu_traffic(wl, "block_lists")
    roof_hbm = {"bound": "hbm", "kernel": "block_lists", "achieved": blk_bytes / (blk_ms * 1e-3) / 1e9 if blk_ms else None,
                "peak": hbm_peak, "unit": "GB/s", "frac": (blk_bytes / (blk_ms * 1e-3) / 1e9) / hbm_peak if blk_ms else None,
                "traffic": blk_traffic, "peak_source": hbm_src, "algorithmic_bytes": blk_bytes,
                "frac_on_dram_traffic": (blk_traffic / (blk_ms * 1e-3) / 1e9) / hbm_peak if (blk_ms and blk_traffic) else None,
                "work": f"{M} tile-list entries x 36 B in + ~{blk_entries} block-list entries x 4 B out",
                "avg_launch_ms": blk_ms,
                "note": "DRAM traffic is above the algorithmic bytes because a 32-byte gather costs a 64-byte DRAM access and records are "
                        "shared by tiles that run far apart; the two binning modes that avoid the gathers (PS_BIN_MODE=bytes|split) remove "
                        "that traffic but lose end to end (DESIGN.md section 7)"}
    per_kernel = per_kernel_traffic(wl, mode, V, args.n or cfg["n"], M, n_lists, stats_all, stage_ms)
    out = {"metric": METRIC, "value": value, "unit": "views/s", "n_gpus": world, "steps": steps, "warmup": max(3, args.warmup),
           "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(wl) + (" -- FORWARD ONLY (inference)" if fwd_only else ""), "mode": mode,
                      "pass": "forward only" if fwd_only else "forward + backward", "views_per_step_per_gpu": V, "frames_per_step_per_gpu": F,
                      "cameras": n_cams, "gaussians_per_frame": args.n or cfg["n"], "isect_per_step": M,
                      "parallelism": f"views sharded over {world} GPU(s), " + ("views split, gradient rows pushed into the owner rank's staging buffer over NVLink by the projection-backward kernel" if fused
                                                                                 else "views split, NCCL all-reduce of d_params" if need_reduce
                                                                                 else "whole frames per rank, no collective"),
                      "l2": f"{n_sets} rotating input batches ({input_mb:.0f} MB) + {V * cfg['n'] * 60 / 2**20:.0f} MB of per-step intermediates > 126 MB L2"},
           "e2e": {"value": e2e_value, "unit": "views/s", "ms_per_step": ms_e2e / steps,
                   "h2d_bytes_per_step": h2d_bytes,
                   "d2h_bytes_per_step": int(V * H * W * 4) if fwd_only else int(gnorm_host[0].numel() * 4 + 4),
                   "result_read_back": "uint8 RGBA images" if fwd_only else
                                       "the step's loss + per-frame gradient norms (d_params stays on the device, where a training step's optimiser consumes it)",
                   "with_gradient_readback": None if ms_e2e_grad is None else {
                       "value": V * world / (ms_e2e_grad * 1e-3), "ms_per_step": ms_e2e_grad,
                       "d2h_bytes_per_step": int(out_host[0].numel() * 4 + gnorm_host[0].numel() * 4 + 4),
                       "what": "round-1 definition: additionally the whole d_params [F,N,P] is copied to pinned host memory every step; "
                               "bounded by the host's device->host rate with all ranks copying at once (copies_alone.d2h_*)"},
                   "h2d_gbs_per_rank": h2d_bytes / (ms_e2e / steps * 1e-3) / 1e9,
                   "copies_alone": {"h2d_ms_per_step": ms_copy / steps, "h2d_gbs_per_rank": h2d_bytes / (ms_copy / steps * 1e-3) / 1e9,
                                    "d2h_ms_per_step": None if ms_copy_out is None else ms_copy_out / steps,
                                    "d2h_gbs_per_rank": None if ms_copy_out is None else out_host[0].numel() * 4 / (ms_copy_out / steps * 1e-3) / 1e9,
                                    "what": "the same copies with no kernels, each direction alone, all ranks at once (max over ranks): what the "
                                            "host side of this box delivers; the e2e step cannot be faster than the slower of these and the kernels"},
                   "host_pinning": getattr(args, "_pin", None),
                   "overlap": "host->device and device->host copies on side streams, overlapped with the neighbouring steps' kernels"},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_hbm": roof_hbm, "binning": binning,
           "stage_ms_per_step": per_step, "pairs": stats_all, "per_kernel": per_kernel, "fp32_peak_tflops": fp32_peak}
    if primary and not fwd_only:
        # the reference's own call shape: ONE view per render() + backward through the drop-in class (SURVEY 8d-d3 iii)
        from pose_splatter_b200 import create_renderer
        r1 = create_renderer(mode, W, H, device=str(dev))
        r1.set_background_color(torch.ones(3))
        p1 = devs[0]["params"][0].clone().requires_grad_(True)
        vm1, K1 = devs[0]["viewmats"][0], devs[0]["Ks"][0]
        lat = []
        for it in range(30):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rgb1, a1 = r1.render(p1, vm1, K1)
            (rgb1.sum() + a1.sum()).backward()
            torch.cuda.synchronize()
            lat.append(1e3 * (time.perf_counter() - t0))
            p1.grad = None
        lat = sorted(lat[5:])
        latf = []
        with torch.no_grad():
            for it in range(30):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                r1.render(p1, vm1, K1)
                torch.cuda.synchronize()
                latf.append(1e3 * (time.perf_counter() - t0))
        latf = sorted(latf[5:])
        out["single_view"] = {"latency_ms_median": lat[len(lat) // 2], "latency_ms_min": lat[0],
                              "forward_only_latency_ms_median": latf[len(latf) // 2], "forward_only_latency_ms_min": latf[0],
                              "what": "one renderer.render(params[N,P], viewmat, K) + backward, host wall clock with a device "
                                      "synchronisation on both sides (launch / latency bound; the batched numbers are the headline)"}
    if world == 1 and primary and not fwd_only and min(H, W) >= 11:
        out["training_step"] = training_step(dev, devs, n_sets, mode, W, H, V, bg, max(3, min(steps, 5)))
    if world == 1 and primary and not fwd_only and mode == "3d":
        out["param_head"] = param_head_leg(dev, F, args.n or cfg["n"])
    if world == 1 and primary and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        probe_vps, _ = cpu_views_per_second(wl, cores, cores, args.n)           # short probe sizes the sample
        sample_views = int(min(4096, max(cores, round(probe_vps * 12.0 / cores) * cores)))  # ~12 s of CPU work
        vps, dt = cpu_views_per_second(wl, sample_views, cores, args.n)
        out["cpu_baseline"] = {"value": vps, "unit": "views/s", "cores": cores, "kind": "port",
                               "sample": f"{sample_views} views of the same workload (full N, full resolution, fwd+bwd), {dt:.1f} s on {cores} threads"}
        # the reference's own pure-PyTorch CPU path (2D only: its 3D arithmetic is gsplat, CUDA-only) beside the port
        out["cpu_baseline_reference"] = reference_2d_cpu("c3" if mode == "3d" else wl)
        # ... and the same unmodified class as torch eager kernels on this GPU
        out["reference_2d_on_b200"] = reference_2d_on_gpu("c3" if mode == "3d" else wl, dev)
    if primary:
        out["gsplat_b200"] = gsplat_comparator()
    if wl == "c4":
        out["full_sequence"] = {"frames": C4_SEQUENCE_FRAMES, "views": C4_SEQUENCE_FRAMES * n_cams,
                                "seconds_at_this_rate_resident": C4_SEQUENCE_FRAMES * n_cams / value,
                                "seconds_at_this_rate_e2e": C4_SEQUENCE_FRAMES * n_cams / e2e_value,
                                "sharding": f"frames round-robin over {world} GPU(s), {F} frames per step per GPU, no collective"}
    return out


def param_head_leg(dev, F, N):
    """MLP output -> render() rows (activations + apply_pose_transform_3d) for the F frames of one step in
    one launch each way; device-timed.  Bytes per row: forward 72 in + 56 out, backward 116 in + 60 out."""
    import torch
    from pose_splatter_b200 import param_head
    g = torch.Generator().manual_seed(5)
    n = F * N
    net = torch.randn(n, 14, generator=g).to(dev).requires_grad_(True)
    probs = (torch.rand(n, generator=g) * 0.7 + 0.27).to(dev)
    grid = (torch.rand(n, 3, generator=g) * 0.2 - 0.1).to(dev)
    cot = torch.randn(n, 14, generator=g).to(dev)
    angles = torch.rand(F, generator=g, dtype=torch.float64) * 6.28 - 3.14
    p3 = torch.rand(F, 3, generator=g) * 0.1 - 0.05
    rf = torch.arange(F).repeat_interleave(N).int().to(dev)
    scale = torch.tensor([-5.5], device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fw = bw = 0.0
    reps = 5
    for it in range(reps + 2):
        net.grad = None
        ev[0].record()
        rows = param_head.gaussian_rows("3d", net, probs, scale, 0.18 / 112, 0.25, grid_sel=grid, angle=angles, p_3d=p3, row_frame=rf)
        ev[1].record()
        rows.backward(cot)
        ev[2].record()
        torch.cuda.synchronize()
        if it >= 2:
            fw += ev[0].elapsed_time(ev[1]) / reps
            bw += ev[1].elapsed_time(ev[2]) / reps
    return {"rows_per_step": n, "frames": F, "forward_ms": fw, "backward_ms": bw,
            "forward_gbs": n * 128 / (fw * 1e-3) / 1e9, "backward_gbs": n * 176 / (bw * 1e-3) / 1e9,
            "what": "ps_param_head_forward / backward through pose_splatter_b200.param_head.gaussian_rows (autograd), all frames of a "
                    "step in one launch; includes the host-side pose table upload and torch's autograd bookkeeping"}


def training_step(dev, devs, n_sets, mode, W, H, V, bg, steps):
    """render -> per-view loss (soft IoU + L1 + SSIM) -> backward, the reference's training step without
    its networks (scripts/training/train_script.py:107-134), with the loss fused into one C-ABI call that also
    emits the renderer's cotangents.  Beside it: the same loss written the reference's way (torch eager ops +
    autograd) on the same B200, on a slice of the views."""
    import torch
    from pose_splatter_b200 import _capi, batched, losses
    g = torch.Generator(device="cpu").manual_seed(11)
    timg = torch.rand(V, 3, H, W, generator=g).to(dev)
    tmask = (torch.rand(V, H, W, generator=g) > 0.5).float().to(dev)
    sl, il = 1.0, 0.5
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    loss_ms = 0.0

    def step(k, timed_loss=False):
        nonlocal loss_ms
        s = devs[k % n_sets]
        rgb, alpha, _, saved = batched.forward_raw(mode, s["params"], s["view_frame"], s["viewmats"], s["Ks"], bg, W, H,
                                                   _capi.FLAG_SAVE_FOR_BACKWARD)
        if timed_loss:
            ev[2].record()
        parts, d_rgb, d_alpha = losses._launch(rgb, alpha, timg, tmask, sl, il, True)
        if timed_loss:
            ev[3].record()
        d_params = batched.backward_raw(saved, s["params"], s["view_frame"], s["viewmats"], s["Ks"], bg, d_rgb, d_alpha)
        saved.release()
        if timed_loss:
            torch.cuda.synchronize()
            loss_ms += ev[2].elapsed_time(ev[3])
        return parts, d_params

    for k in range(3):
        step(k)
    torch.cuda.synchronize()
    ev[0].record()
    for k in range(steps):
        step(k)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    for k in range(steps):
        step(k, timed_loss=True)
    loss_ms /= steps

    # the reference's formulation on the same GPU: three eager torch graphs + autograd (torchmetrics' SSIM restated)
    Vc = min(V, 96)
    s = devs[0]
    with torch.no_grad():
        rgb, alpha, _, _ = batched.forward_raw(mode, s["params"], s["view_frame"], s["viewmats"], s["Ks"], bg, W, H, 0)
    rgb, alpha = rgb[:Vc].clone(), alpha[:Vc].clone()
    d = torch.arange(-5, 6, dtype=torch.float32, device=dev)
    taps = torch.exp(-((d / 1.5) ** 2) / 2)
    taps = taps / taps.sum()
    k2d = torch.outer(taps, taps)[None, None].expand(3, 1, 11, 11).contiguous()

    def eager():
        r = rgb.requires_grad_(True)
        a = alpha.requires_grad_(True)
        q = r.permute(0, 3, 1, 2)
        t, m = timg[:Vc], tmask[:Vc]
        inter = (a * m).sum(dim=(-2, -1))
        union = (a + m - a * m).sum(dim=(-2, -1))
        iou = 1 - (inter + 1e-6) / (union + 1e-6)
        stack = torch.cat([t, q, t * t, q * q, t * q])
        o = torch.nn.functional.conv2d(stack, k2d, groups=3)
        mp, mq, epp, eqq, epq = o.split(Vc)
        spp, sqq, spq = torch.clamp(epp - mp * mp, min=0), torch.clamp(eqq - mq * mq, min=0), epq - mp * mq
        smap = ((2 * mp * mq + 1e-4) * (2 * spq + 9e-4)) / ((mp * mp + mq * mq + 1e-4) * (spp + sqq + 9e-4))
        ssim = sl * (1 - smap.mean(dim=(1, 2, 3)))
        img = il * (t - q).abs().sum(dim=(1, 2, 3)) / m.sum(dim=(-2, -1))
        (iou + ssim + img).sum().backward()
        r.grad = a.grad = None

    for _ in range(2):
        eager()
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(3):
        eager()
    ev[1].record()
    torch.cuda.synchronize()
    eager_ms = ev[0].elapsed_time(ev[1]) / 3 * (V / Vc)
    return {"value": V / (ms * 1e-3), "unit": "views/s", "ms_per_step": ms, "views_per_step": V,
            "loss_ms_per_step": loss_ms, "loss_algorithmic_gbs": V * H * W * 4 * (3 + 1 + 3 + 1 + 3 + 1) / (loss_ms * 1e-3) / 1e9,
            "torch_eager_loss_ms_per_step": eager_ms,
            "what": "render forward -> ps_view_loss (soft IoU + 0.5 * L1 / sum(mask) + 1.0 * (1 - SSIM), losses and "
                    "d_rgb / d_alpha in one call) -> render backward, inputs resident; torch_eager = the same loss as "
                    f"eager torch ops + autograd on this GPU, measured on {Vc} views and scaled to {V}"}


if __name__ == "__main__":
    a = parse()
    if a.workload == "c4":
        a.forward_only = True  # configs[3] is inference: frames of the sequence sharded across ranks, no backward
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
