"""On-disk formats either side of the renderer.

* Gaussian NPZ -- the only persisted form of real `gaussian_params` in the reference
  (scripts/visualization/export_gaussian_full.py:113-137 collects, :163-178 `save_npz` writes): activated values
  `means` (centred), `quaternions` (wxyz, as the network emits them: not normalised), `scales` (= exp(log_scales)),
  `opacities` (= sigmoid(logit), [N]), `colors`, `center` [1,3], `metadata`.  `save_gaussian_npz` writes the same
  archive; `load_gaussian_npz` reads it back into the [N,14] ACTIVATED row layout that
  `render_views(..., activated=True)` / PS_FLAG_ACTIVATED_INPUTS takes, so an exported frame replays through the
  renderer (and the benchmark) without the model.
* uint8 RGBA evaluation renders: written on the GPU by `render_views_rgba8` (scripts/utils/evaluate_model.py:101-113) and
  taken to disk by `RenderSequenceWriter`, the sink of the reference's evaluation loop (:78-146): dataset `images`
  [T,C,h,w,4] uint8, written in slabs of `write_batch_frames` = 50 frames.  The reference builds every slab from a
  synchronous `.cpu().numpy()` per frame; here a slab leaves the GPU as ONE asynchronous copy into pinned memory on a side
  stream and a writer thread stores it while the next frames render.  Container: gzip HDF5 exactly like the reference when
  h5py is importable, otherwise (h5py is absent from this image) a plain `.npy` memory map with the same array.

Host-side only: no kernel work here.
"""
from __future__ import annotations

import queue
import threading

import numpy as np
import torch

NPZ_KEYS = ("means", "quaternions", "scales", "opacities", "colors", "center")


def activate_rows(gaussian_params: torch.Tensor):
    """Raw [N,14] rows -> (means, quats, scales, opacities, colors) exactly as the reference's legacy accessor does
    (src/model.py:300-317): scales = exp(log_scales), opacities = sigmoid(logit)[N]; quats and colours untouched."""
    if gaussian_params.dim() != 2 or gaussian_params.shape[1] != 14:
        raise ValueError(f"Expected 14 parameters per Gaussian, got {tuple(gaussian_params.shape)}")
    p = gaussian_params.detach().float().cpu()
    return p[:, 0:3], p[:, 6:10], torch.exp(p[:, 3:6]), torch.sigmoid(p[:, 13]), p[:, 10:13]


def save_gaussian_npz(filename, means, quats, scales, opacities, colors, center=None):
    """Same archive as the reference's `save_npz` (export_gaussian_full.py:163-178).  With center=None the point
    cloud is centred here like :126-128 (center = mean of the means, subtracted before saving)."""
    means = np.asarray(means, np.float32)
    if center is None:
        center = means.mean(axis=0, keepdims=True)
        means = means - center
    np.savez_compressed(filename, means=means, quaternions=np.asarray(quats, np.float32),
                        scales=np.asarray(scales, np.float32), opacities=np.asarray(opacities, np.float32),
                        colors=np.asarray(colors, np.float32), center=np.asarray(center, np.float32),
                        metadata={"format": "gaussian_splatting_full", "num_gaussians": len(means), "version": "1.0"})


def load_gaussian_npz(filename) -> torch.Tensor:
    """NPZ -> [N,14] activated rows: means + center | scales | quats | colours | opacity."""
    # allow_pickle stays off: every array this loader needs is plain float data.  The reference's `metadata` entry is a
    # pickled dict (export_gaussian_full.py:176); unpickling an archive someone else wrote can run arbitrary code, so it
    # is not read -- the required keys identify the format.
    with np.load(filename, allow_pickle=False) as z:
        missing = [k for k in NPZ_KEYS if k not in z.files]
        if missing:
            raise ValueError(f"{filename}: not a gaussian_splatting_full archive (missing {missing})")
        means = z["means"].astype(np.float32) + z["center"].astype(np.float32).reshape(1, 3)
        n = len(means)
        cols = [means, z["scales"].astype(np.float32).reshape(n, 3), z["quaternions"].astype(np.float32).reshape(n, 4),
                z["colors"].astype(np.float32).reshape(n, 3), z["opacities"].astype(np.float32).reshape(n, 1)]
    return torch.from_numpy(np.concatenate(cols, axis=1))


class RenderSequenceWriter:
    """uint8 RGBA renders of a frame sequence -> `images` [T,C,h,w,4] on disk, in slabs of `write_batch_frames` frames
    (scripts/utils/evaluate_model.py:78-146: `write_batch_frames = 50`, gzip level `image_compression_level`).

        with RenderSequenceWriter(fn, T, C, h, w) as out:
            for f0 in range(0, T, B):
                out.put(render_views_rgba8("3d", params[f0:f0 + B], view_frame, w, h, bg, viewmats, Ks))   # [B*C,h,w,4]

    `put` accepts [F,C,h,w,4] or [F*C,h,w,4] uint8 tensors (CUDA or host) and returns as soon as the device->host copies are
    queued: a slab is copied into one of `n_buffers` pinned host buffers on a side stream (ordered behind the producer's
    current stream by an event) and written by a background thread; `put` blocks only when every buffer is still on its
    way to disk.  Frames are appended in call order from `first_frame` (a rank of a sharded run passes the start of its
    own frame range and its own file, or shares one `.npy` map: disjoint ranges of a memory map may be written by several
    processes).  `close()` flushes; errors of the writer thread surface in the next `put` / `close`."""

    def __init__(self, filename, total_frames: int, n_cams: int, height: int, width: int, write_batch_frames: int = 50,
                 compression_level: int = 4, first_frame: int = 0, backend: str | None = None, n_buffers: int = 3,
                 create: bool = True):
        if min(total_frames, n_cams, height, width, write_batch_frames, n_buffers) <= 0:
            raise ValueError("RenderSequenceWriter: sizes must be positive")
        self.shape = (int(total_frames), int(n_cams), int(height), int(width), 4)
        self.slab = int(write_batch_frames)
        self.next_frame = int(first_frame)
        if backend is None:
            backend = "npy"
            if str(filename).endswith((".h5", ".hdf5")):
                try:
                    import h5py  # noqa: F401
                    backend = "h5py"
                except ImportError as e:
                    raise RuntimeError(f"{filename}: h5py is not installed; use a .npy file name (same array, memory-mapped)") from e
        self.backend = backend
        if backend == "h5py":
            import h5py
            self._file = h5py.File(str(filename), "w" if create else "r+")
            self._data = (self._file.create_dataset("images", self.shape, dtype="uint8", compression="gzip",
                                                    compression_opts=int(compression_level)) if create else self._file["images"])
        elif backend == "npy":
            self._file = None
            self._data = np.lib.format.open_memmap(str(filename), mode="w+" if create else "r+", dtype=np.uint8,
                                                   shape=self.shape if create else None)
            if tuple(self._data.shape) != self.shape:
                raise ValueError(f"{filename}: holds {tuple(self._data.shape)}, expected {self.shape}")
        else:
            raise ValueError(f"RenderSequenceWriter: unknown backend '{backend}'")
        self._free: queue.Queue = queue.Queue()
        self._todo: queue.Queue = queue.Queue()
        self._n_buffers = int(n_buffers)
        self._made = 0
        self._pinned = None  # decided by the first put
        self._copy_stream = None
        self._error = None
        self._closed = False
        self._thread = threading.Thread(target=self._drain, name="RenderSequenceWriter", daemon=True)
        self._thread.start()

    # -- writer thread: wait for the slab's copy, store it, hand the buffer back
    def _drain(self):
        while True:
            item = self._todo.get()
            if item is None:
                return
            buf, n, i1, event = item
            try:
                if self._error is None:
                    if event is not None:
                        event.synchronize()
                    self._data[i1:i1 + n] = buf[:n].numpy()
            except Exception as e:  # noqa: BLE001 -- re-raised on the caller's thread
                self._error = e
            finally:
                self._free.put(buf)

    def _buffer(self, pinned: bool):
        if self._free.empty() and self._made < self._n_buffers:
            self._made += 1
            return torch.empty((self.slab,) + self.shape[1:], dtype=torch.uint8, pin_memory=pinned)
        return self._free.get()  # blocks while every buffer is still being written

    def _check(self):
        if self._error is not None:
            e, self._error = self._error, None
            raise RuntimeError(f"RenderSequenceWriter: the writer thread failed: {e}") from e

    def put(self, rgba8: torch.Tensor):
        if self._closed:
            raise RuntimeError("RenderSequenceWriter: put() after close()")
        self._check()
        C, h, w = self.shape[1:4]
        if rgba8.dtype != torch.uint8:
            raise ValueError(f"Expected uint8 RGBA, got {rgba8.dtype}")
        if rgba8.dim() == 4 and tuple(rgba8.shape[1:]) == (h, w, 4) and rgba8.shape[0] % C == 0:
            rgba8 = rgba8.reshape(rgba8.shape[0] // C, C, h, w, 4)
        if rgba8.dim() != 5 or tuple(rgba8.shape[1:]) != (C, h, w, 4):
            raise ValueError(f"Expected [F,{C},{h},{w},4] or [F*{C},{h},{w},4], got {tuple(rgba8.shape)}")
        F = int(rgba8.shape[0])
        if self.next_frame + F > self.shape[0]:
            raise ValueError(f"RenderSequenceWriter: frames {self.next_frame}..{self.next_frame + F} exceed the {self.shape[0]} of the file")
        rgba8 = rgba8.contiguous()
        on_gpu = rgba8.is_cuda
        if self._pinned is None:
            self._pinned = on_gpu
        ready = None
        if on_gpu:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=rgba8.device)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(rgba8.device))  # the frames are complete behind this point
        for f0 in range(0, F, self.slab):
            n = min(self.slab, F - f0)
            buf = self._buffer(self._pinned)
            event = None
            if on_gpu:
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(ready)
                    buf[:n].copy_(rgba8[f0:f0 + n], non_blocking=True)
                    event = torch.cuda.Event()
                    event.record(self._copy_stream)
                rgba8.record_stream(self._copy_stream)  # the allocator must not recycle the frames under the copy
            else:
                buf[:n].copy_(rgba8[f0:f0 + n])
            self._todo.put((buf, n, self.next_frame + f0, event))
        self.next_frame += F

    def close(self):
        if self._closed:
            return
        self._closed = True
        self._todo.put(None)
        self._thread.join()
        if self.backend == "npy":
            self._data.flush()
        if self._file is not None:
            self._file.close()
        self._data = None
        self._check()

    def __del__(self):  # a writer dropped without close() still flushes what it was given
        try:
            self.close()
        except Exception:  # noqa: BLE001 -- nothing to report to at collection time
            pass

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        self.close()
        return False


def load_render_sequence(filename):
    """The array a RenderSequenceWriter wrote: [T,C,h,w,4] uint8 (HDF5 dataset `images`, or the .npy memory map)."""
    if str(filename).endswith((".h5", ".hdf5")):
        import h5py
        with h5py.File(str(filename), "r") as f:
            return f["images"][...]
    return np.load(str(filename), mmap_mode="r")
