"""Drop-in replacement for the reference's ``cascade_model.py`` (``cfg.CASCADE_MODEL``; selected by
reference trainer.py:83-84): the CP-CSV generator whose image trunk is modulated by a RE-ENCODING of
its own segmentation mask (``presample`` + four ``downBlock`` s) instead of the segmentation
trunk's activations, plus ``train_autoencoder``; the discriminators are those of ``model.py``.

Same surface as the reference file: ``StoryGAN.sample_videos`` / ``sample_images`` return
``((zmc_seg, h_seg1, h_seg2, h_seg3), (g_seg1, g_seg2, g_seg3, g_seg4))`` in slot 0
(cascade_model.py:441-445, 514-518), ``train_autoencoder(real_segments)`` returns the
reconstructed mask (cascade_model.py:528-540), ``downBlock`` / ``presample`` parameter holders
keep the reference's state-dict keys.  The arithmetic runs on the kernel tapes of
``cpcsv_b200.cascade``; there is no PyTorch or CPU fallback.

Like ``model.py`` this file may be copied next to a training run and re-imported under another
name (reference trainer.py:55-61 copies it to ``<output_dir>/model.py``): the base classes are
loaded from the installed ``model.py`` by path (``CPCSV_B200_HOME`` or next to ``cpcsv_b200``).
"""
import importlib.util
import os
import sys

import torch
import torch.nn as nn

try:
    import cpcsv_b200  # noqa: F401
except ImportError:
    sys.path.insert(0, os.environ.get("CPCSV_B200_HOME", os.path.dirname(os.path.abspath(__file__))))
    import cpcsv_b200  # noqa: F401

from cpcsv_b200 import cascade as kcascade


def _base_model():
    home = os.environ.get("CPCSV_B200_HOME") or os.path.dirname(os.path.dirname(os.path.abspath(cpcsv_b200.__file__)))
    path = os.path.join(home, "model.py")
    mod = sys.modules.get("model")
    if mod is not None and os.path.abspath(getattr(mod, "__file__", "")) == os.path.abspath(path) \
            and hasattr(mod, "STAGE1_D_STY_V2"):
        return mod
    name = "_cpcsv_b200_base_model"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_base = _base_model()
cfg = _base.cfg
conv3x3, upBlock = _base.conv3x3, _base.upBlock
CA_NET, D_GET_LOGITS = _base.CA_NET, _base.D_GET_LOGITS
STAGE1_D_IMG, STAGE1_D_SEG, STAGE1_D_STY_V2 = _base.STAGE1_D_IMG, _base.STAGE1_D_SEG, _base.STAGE1_D_STY_V2


def downBlock(in_planes, out_planes):
    """conv3x3 stride 2 pad 1 (with bias) -> BatchNorm2d -> ReLU holder (reference cascade_model.py:36-41)."""
    return nn.Sequential(nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=2, padding=1),
                         nn.BatchNorm2d(out_planes), nn.ReLU(True))


class StoryGAN(_base.StoryGAN):
    """Cascade CP-CSV generator (reference cascade_model.py:221-540)."""

    def define_module(self):
        super(StoryGAN, self).define_module()
        nseg = self.gf_dim_seg
        self.presample = nn.Sequential(conv3x3(1, nseg // 16), nn.BatchNorm2d(nseg // 16), nn.ReLU())
        self.downsample1_seg = downBlock(nseg // 16, nseg // 8)
        self.downsample2_seg = downBlock(nseg // 8, nseg // 4)
        self.downsample3_seg = downBlock(nseg // 4, nseg // 2)
        self.downsample4_seg = downBlock(nseg // 2, nseg)

    def _needs_grad(self, x):
        return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))

    def _trunk(self, zmc_all, seg):
        outs = kcascade.CascadeTrunkRunner(self, self._needs_grad(zmc_all), seg).apply(zmc_all)
        img, segm = outs[0], outs[1]
        return (tuple(outs[2:6]), tuple(outs[6:10])), img, segm

    def train_autoencoder(self, real_segments):
        return kcascade.AutoencoderRunner(self, self._needs_grad(real_segments)).apply(real_segments)
