"""the reference's Gaussian NPZ archive (scripts/visualization/export_gaussian_full.py:163-178) written,
read back and replayed through the renderer."""
import numpy as np
import pytest
import torch


def _rows(n=300, seed=0):
    from pose_splatter_b200 import synth
    return synth.gaussians_3d(n, seed)


def test_npz_roundtrip_matches_reference_archive_layout(tmp_path):
    from pose_splatter_b200 import formats
    p = _rows()
    means, quats, scales, opac, colors = formats.activate_rows(p)
    fn = tmp_path / "gaussian_frame0000.npz"
    formats.save_gaussian_npz(fn, means.numpy(), quats.numpy(), scales.numpy(), opac.numpy(), colors.numpy())
    z = np.load(fn, allow_pickle=True)
    assert sorted(z.files) == sorted(["means", "quaternions", "scales", "opacities", "colors", "center", "metadata"])
    assert z["center"].shape == (1, 3) and abs(z["means"].mean(0)).max() < 1e-6          # centred like :126-128
    assert z["metadata"].item() == {"format": "gaussian_splatting_full", "num_gaussians": 300, "version": "1.0"}
    assert z["opacities"].shape == (300,)
    rows = formats.load_gaussian_npz(fn)
    assert rows.shape == (300, 14)
    assert torch.allclose(rows[:, 0:3], means, atol=1e-6)
    assert torch.equal(rows[:, 3:6], scales) and torch.equal(rows[:, 6:10], quats)
    assert torch.equal(rows[:, 10:13], colors) and torch.equal(rows[:, 13], opac)


def test_npz_errors(tmp_path):
    from pose_splatter_b200 import formats
    with pytest.raises(ValueError, match="Expected 14 parameters"):
        formats.activate_rows(torch.zeros(5, 9))
    np.savez(tmp_path / "bad.npz", means=np.zeros((2, 3)))
    with pytest.raises(ValueError, match="missing"):
        formats.load_gaussian_npz(tmp_path / "bad.npz")


@pytest.mark.gpu
def test_npz_replay_renders_the_same_image(tmp_path):
    """raw rows rendered directly == the exported archive rendered as activated inputs (colours already in [0,1])"""
    from pose_splatter_b200 import batched, formats, synth
    d = synth.make_views("c2", n_frames=1, n_cams=2, seed=3, n=800)
    W, H = d["width"], d["height"]
    dev = torch.device("cuda")
    bg = torch.ones(3, device=dev)
    args = (d["view_frame"].to(dev), W, H, bg, d["viewmats"].to(dev), d["Ks"].to(dev))
    rgb0, a0 = batched.render_views("3d", d["params"].to(dev), *args)
    means, quats, scales, opac, colors = formats.activate_rows(d["params"][0])
    fn = tmp_path / "frame.npz"
    formats.save_gaussian_npz(fn, means.numpy(), quats.numpy(), scales.numpy(), opac.numpy(), colors.numpy())
    rows = formats.load_gaussian_npz(fn)[None].to(dev)
    rgb1, a1 = batched.render_views("3d", rows, *args, activated=True)
    # the archive stores centred means: adding the centre back is not bit-exact, and quats are normalised with
    # q * rsqrt(q.q) instead of q / (|q| + 1e-8) on this path
    assert (rgb0 - rgb1).abs().max().item() < 1e-2 and (rgb0 - rgb1).abs().mean().item() < 1e-5  # a 1/255 alpha-threshold flip is 4e-3
    assert (a0 - a1).abs().max().item() < 1e-2 and (a0 - a1).abs().mean().item() < 1e-5


# ---- RenderSequenceWriter (host logic; the device path differs only by the pinned buffers and the copy stream) ----------
def _frames(T, C, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (T, C, h, w, 4), dtype=torch.uint8, generator=g)


def test_render_sequence_writer_slabs_and_ragged_puts(tmp_path):
    from pose_splatter_b200 import formats
    T, C, h, w = 137, 3, 9, 7
    frames = _frames(T, C, h, w)
    fn = tmp_path / "renders.npy"
    with formats.RenderSequenceWriter(fn, T, C, h, w, write_batch_frames=50, n_buffers=2) as out:
        f = 0
        for n in (1, 49, 3, 60, 24):                       # slab boundaries crossed, partial last slab
            piece = frames[f:f + n]
            out.put(piece if n % 2 else piece.reshape(n * C, h, w, 4))   # both accepted layouts
            f += n
        assert f == T and out.next_frame == T
    back = formats.load_render_sequence(fn)
    assert back.shape == (T, C, h, w, 4) and back.dtype == np.uint8
    assert np.array_equal(np.asarray(back), frames.numpy())


def test_render_sequence_writer_ranges_of_two_ranks_share_one_map(tmp_path):
    from pose_splatter_b200 import formats
    T, C, h, w = 40, 2, 5, 6
    frames = _frames(T, C, h, w, seed=1)
    fn = tmp_path / "renders.npy"
    a = formats.RenderSequenceWriter(fn, T, C, h, w, write_batch_frames=8, first_frame=0)
    b = formats.RenderSequenceWriter(fn, T, C, h, w, write_batch_frames=8, first_frame=25, create=False)
    b.put(frames[25:40])
    a.put(frames[0:25])
    a.close()
    b.close()
    assert np.array_equal(np.asarray(formats.load_render_sequence(fn)), frames.numpy())


def test_render_sequence_writer_errors(tmp_path):
    from pose_splatter_b200 import formats
    out = formats.RenderSequenceWriter(tmp_path / "r.npy", 4, 2, 3, 3)
    with pytest.raises(ValueError, match="uint8"):
        out.put(torch.zeros(1, 2, 3, 3, 4))
    with pytest.raises(ValueError, match="Expected"):
        out.put(torch.zeros(1, 2, 3, 5, 4, dtype=torch.uint8))
    out.put(torch.zeros(3, 2, 3, 3, 4, dtype=torch.uint8))
    with pytest.raises(ValueError, match="exceed"):
        out.put(torch.zeros(2, 2, 3, 3, 4, dtype=torch.uint8))
    out.close()
    with pytest.raises(RuntimeError, match="after close"):
        out.put(torch.zeros(1, 2, 3, 3, 4, dtype=torch.uint8))
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(RuntimeError, match="h5py is not installed"):
            formats.RenderSequenceWriter(tmp_path / "r.h5", 4, 2, 3, 3)
    with pytest.raises(ValueError, match="positive"):
        formats.RenderSequenceWriter(tmp_path / "z.npy", 0, 2, 3, 3)
