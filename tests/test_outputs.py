"""miscc/outputs.py (the consumers of the generator's outputs, re-exported from miscc.utils under the
reference's names) against torchvision.utils, which is what the reference builds them from
(reference miscc/utils.py:205-311, 343-428)."""
import os
import types

import numpy as np
import pytest
import torch

vutils = pytest.importorskip("torchvision.utils")

from miscc.config import cfg  # noqa: E402
import miscc.utils as mu  # noqa: E402
from miscc import outputs  # noqa: E402


def _cfg(video_len=5, st_batch=3):
    cfg.VIDEO_LEN, cfg.TRAIN.ST_BATCH_SIZE, cfg.IMSIZE, cfg.TEXT.DIMENSION = video_len, st_batch, 64, 7


@pytest.mark.parametrize("shape,per_row", [((5, 3, 8, 6), 5), ((7, 3, 8, 6), 3), ((4, 1, 5, 5), 2), ((1, 3, 4, 4), 8)])
def test_sheet_matches_make_grid(shape, per_row):
    x = torch.randn(*shape)
    assert torch.equal(outputs._sheet(x, per_row), vutils.make_grid(x, per_row))
    assert torch.equal(outputs._sheet(list(x), per_row), vutils.make_grid(list(x), per_row))


def test_story_and_image_results_match_reference_construction(tmp_path):
    _cfg()
    B, V = 3, 5
    fake, real = torch.rand(B, 3, V, 64, 64) * 2.4 - 1.2, torch.rand(B, 3, V, 64, 64) * 2 - 1

    def ref_numpy(t):       # reference images_to_numpy
        g = t.numpy().transpose(1, 2, 0).copy()
        g[g < -1], g[g > 1] = -1, 1
        return ((g + 1) / 2 * 255).astype("uint8")

    def ref_sheet(stories):
        rows = [vutils.make_grid(torch.transpose(stories[i], 0, 1), V) for i in range(B)]
        return ref_numpy(vutils.make_grid(rows, 1))
    texts = [["frame %d of story %d" % (t, b) for b in range(B)] for t in range(V)]
    got = mu.save_story_results(real, fake, texts, "007", str(tmp_path))
    assert np.array_equal(got, np.concatenate([ref_sheet(fake), ref_sheet(real)], axis=1))
    lines = open(tmp_path / "fake_samples_007.txt").read().split("\n")
    assert lines[0].startswith("0---") and lines[1] == "frame 0 of story 0" and "frame 4 of story 2" in lines
    seg = torch.rand(B * V, 1, 64, 64) * 2 - 1
    got = mu.save_image_results(None, seg)
    r = seg.reshape(B, V, -1, 64, 64)
    ref = ref_numpy(vutils.make_grid([vutils.make_grid(r[i], V) for i in range(B)], 1))
    assert np.array_equal(got, ref)


def test_save_all_img_matches_torchvision_save_image(tmp_path):
    import PIL.Image
    images = torch.rand(2, 3, 3, 16, 16) * 2 - 1            # [-1, 1]: negative pixels clip to black, as in the reference
    a, b = tmp_path / "mine", tmp_path / "tv"
    a.mkdir(), b.mkdir()
    assert mu.save_all_img(images, 10, str(a)) == 16
    k = 10
    for s in range(2):
        for i in range(3):
            k += 1
            vutils.save_image(images[s].transpose(0, 1)[i], str(b / ("%d.png" % k)))
            assert np.array_equal(np.asarray(PIL.Image.open(a / ("%d.png" % k))),
                                  np.asarray(PIL.Image.open(b / ("%d.png" % k))))


def test_sampling_loops_write_the_reference_files(tmp_path, monkeypatch):
    _cfg(video_len=3, st_batch=2)
    monkeypatch.chdir(tmp_path)
    calls = []

    class FakeG(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.zeros(1))

        def sample_videos(self, motion_input, content_input, seg=False):
            calls.append((tuple(motion_input.shape), tuple(content_input.shape)))
            return None, torch.rand(motion_input.shape[0], 3, 3, 64, 64) * 2 - 1, None, None, None, None, None
    loader = [dict(images=torch.rand(2, 3, 3, 64, 64), description=torch.randn(2, 3, 9), labels=torch.ones(2, 3, 4),
                   text=[["a", "b"]] * 3) for _ in range(2)]
    out = tmp_path / "Test"
    out.mkdir()
    mu.save_test_samples(FakeG(), loader, str(out))
    assert calls[0] == ((2, 3, 7 + 4), (2, 3, 7))            # text cut to cfg.TEXT.DIMENSION, labels appended
    assert np.load(out / "images.npy").shape == (4, 3, 3, 64, 64) and np.load(out / "labels.npy").shape == (4, 3, 4)
    assert {"fake_samples_000.txt", "fake_samples_001.txt"} <= set(os.listdir(out))
    mu.inference_samples(FakeG(), loader, str(tmp_path / "gen"))
    assert len(os.listdir(tmp_path / "gen")) == 12 and len(os.listdir(tmp_path / "Evaluation" / "ref")) == 12
    stories, labels = mu.create_random_shuffle(torch.randn(6, 3, 5, 8, 8), random_rate=0.5)
    assert stories.shape == (6, 3, 5, 8, 8) and labels.shape == (6,) and set(labels.tolist()) <= {0.0, 1.0}
    assert isinstance(types.FunctionType, type) and mu.check_is_order([1, 2, 2]) and not mu.check_is_order([2, 1])
