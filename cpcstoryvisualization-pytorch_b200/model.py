"""Drop-in replacement for the reference's ``model.py`` (CP-CSV generator and the image /
story / segmentation discriminators) whose arithmetic runs in hand-written sm_100a kernels
(libcpcsv.so through ``cpcsv_b200``).

Same public surface as the reference module: class names, constructor
arguments, ``sample_videos`` / ``sample_images`` 7-tuples, discriminator attributes
(``get_cond_logits``, ``get_uncond_logits``, ``cate_classify``, ``seq_consisten_model``),
``state_dict`` keys incl. the legacy spectral-norm ``weight_orig / weight_u / weight_v`` and
BatchNorm buffers, and parameter-holder class names (``Conv*``, ``BatchNorm*``, ``Linear``)
so the reference's ``weights_init`` keeps working.  The nn.Conv2d / nn.BatchNorm2d /
nn.Linear / nn.GRUCell instances below only HOLD parameters; they are never called -- the
compute is dispatched to the kernel tapes in ``cpcsv_b200.nets``.  There is no PyTorch or CPU
fallback: without a CUDA device and the built extension every forward raises.

This file may be copied next to a training run and re-imported under another name
(reference trainer.py:55-61, inference.py:61-68); it therefore uses absolute imports only and
locates ``cpcsv_b200`` through ``sys.path`` or the ``CPCSV_B200_HOME`` environment variable.
"""
import os
import sys

import torch
import torch.nn as nn
from torch.nn.utils import spectral_norm

try:
    import cpcsv_b200  # noqa: F401
except ImportError:  # copied elsewhere: fall back to the recorded install location
    _home = os.environ.get("CPCSV_B200_HOME", os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, _home)
    import cpcsv_b200  # noqa: F401

from cpcsv_b200 import functions as Fx
from cpcsv_b200 import nets, streams
from miscc.config import cfg


# --------------------------------------------------------------------------- building blocks
def conv3x3(in_planes, out_planes, stride=1, use_spectral_norm=False):
    """3x3 convolution, padding 1, no bias (parameter holder; reference model.py:16-22)."""
    conv = nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)
    return spectral_norm(conv) if use_spectral_norm else conv


def upBlock(in_planes, out_planes):
    """nearest x2 -> conv3x3 -> BatchNorm2d -> ReLU holder with the reference's child indices
    (conv at 1, BN at 2; reference model.py:26-34)."""
    return nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"), conv3x3(in_planes, out_planes),
                         nn.BatchNorm2d(out_planes), nn.ReLU(True))


def _linear_bn(n_in, n_out, bias=True, tail=None):
    mods = [nn.Linear(n_in, n_out, bias=bias), nn.BatchNorm1d(n_out)]
    if tail is not None:
        mods.append(tail)
    return nn.Sequential(*mods)


class CA_NET(nn.Module):
    """Conditioning augmentation (reference model.py:37-65)."""

    def __init__(self):
        super(CA_NET, self).__init__()
        self.t_dim = cfg.TEXT.DIMENSION * cfg.VIDEO_LEN
        self.c_dim = cfg.GAN.CONDITION_DIM
        self.fc = nn.Linear(self.t_dim, self.c_dim * 2, bias=True)
        self.relu = nn.ReLU()

    def draw_eps(self, like):
        """N(0,1) noise for the reparameterisation; tests replace this to inject noise."""
        return torch.randn(like.shape, device=like.device, dtype=like.dtype)

    def forward(self, text_embedding):
        pre = Fx.linear(text_embedding, self.fc.weight, self.fc.bias)
        eps = self.draw_eps(pre[:, :self.c_dim])
        mu, logvar, c_code = Fx.CondAugFn.apply(pre, eps)
        return c_code, mu, logvar

    def encode(self, text_embedding):
        _, mu, logvar = self.forward(text_embedding)
        return mu, logvar


class D_GET_LOGITS(nn.Module):
    """Conditional logits head (reference model.py:68-97)."""

    def __init__(self, ndf, nef, bcondition=True):
        super(D_GET_LOGITS, self).__init__()
        self.df_dim, self.ef_dim, self.bcondition = ndf, nef, bcondition
        if not bcondition:
            raise NotImplementedError("unconditional logits are unused by CP-CSV (get_uncond_logits is None)")
        self.outlogits = nn.Sequential(
            conv3x3(ndf * 8 + nef, ndf * 8, use_spectral_norm=True),
            nn.BatchNorm2d(ndf * 8),
            nn.LeakyReLU(0.2, inplace=True),
            spectral_norm(nn.Conv2d(ndf * 8, 1, kernel_size=4, stride=4)),
            nn.Sigmoid())

    def forward(self, h_code, c_code=None):
        if c_code is None:
            raise NotImplementedError("D_GET_LOGITS needs the condition vector on this path")
        if torch.is_grad_enabled() and c_code.requires_grad:
            raise RuntimeError("cpcsv_b200: D_GET_LOGITS does not return a gradient for c_code (the reference always "
                               "passes a detached condition, miscc/utils.py:56,130); detach it")
        c_code = c_code.reshape(-1, self.ef_dim)
        need_grad = torch.is_grad_enabled() and (
            h_code.requires_grad or any(p.requires_grad for p in self.parameters()))
        return nets.LogitsRunner(self, need_grad).apply(h_code, c_code)


class CateClassifyConv2d(nn.Conv2d):
    """``nn.Conv2d(ndf*8, label_num, 4, 4, 1, bias=False)`` on the 4x4 feature map
    (reference model.py:520): one output pixel, i.e. a dot product per class."""

    def forward(self, h_code):
        n, C = h_code.shape[0], h_code.shape[1]
        flat = h_code.permute(0, 2, 3, 1).reshape(n, 16 * C)        # NHWC-flattened features
        Cp = C
        rows = nets.cate_classify_weight_rows(self.weight, Cp)
        return Fx.linear(flat, rows, None).view(n, self.out_channels, 1, 1)


# --------------------------------------------------------------------------- order-consistency critic
class R2Plus1dStem(nn.Sequential):
    """Parameter holders of the reference's stem (model.py:99-113): spectral-norm Conv3d (1, 7, 7) stride
    (1, 2, 2) -> BatchNorm3d -> ReLU -> spectral-norm Conv3d (1, 1, 1) with temporal padding 1 -> BatchNorm3d
    -> ReLU."""

    def __init__(self):
        super(R2Plus1dStem, self).__init__(
            spectral_norm(nn.Conv3d(3, 45, kernel_size=(1, 7, 7), stride=(1, 2, 2), padding=(0, 3, 3), bias=False)),
            nn.BatchNorm3d(45),
            nn.ReLU(inplace=True),
            spectral_norm(nn.Conv3d(45, 64, kernel_size=(1, 1, 1), stride=(1, 1, 1), padding=(1, 0, 0), bias=False)),
            nn.BatchNorm3d(64),
            nn.ReLU(inplace=True))


class BasicBlock(nn.Module):
    """Importable name of the reference (model.py:115-148): the residual R(2+1)D block.  The reference defines it
    and never instantiates it (``VideoEncoder`` is a plain chain), so it exists here for API compatibility only:
    it holds the same children under the same names and has no kernel path."""
    expansion = 1

    def __init__(self, inplanes, planes, conv_builder, stride=1, downsample=None):
        super(BasicBlock, self).__init__()
        mid = (inplanes * planes * 27) // (inplanes * 9 + 3 * planes)
        self.conv1 = nn.Sequential(conv_builder(inplanes, planes, mid, stride), nn.BatchNorm3d(planes),
                                   nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(conv_builder(planes, planes, mid), nn.BatchNorm3d(planes))
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        raise NotImplementedError("BasicBlock is never built by the reference's networks (model.py:150-197 uses a "
                                  "plain Conv3d chain); there is no accelerated path for it")


class VideoEncoder(nn.Module):
    """Order-consistency critic (reference model.py:150-210): story [B, 3, T, H, W] -> one logit per story.
    The modules below only hold parameters (same state_dict keys as the reference); the convolutions run on the
    kernel tape of ``cpcsv_b200.video``, the detector in the fp32 kernels of ``cpcsv_b200.functions``."""

    def __init__(self):
        super(VideoEncoder, self).__init__()

        def spatial(cin, cout):
            return spectral_norm(nn.Conv3d(cin, cout, kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1),
                                           bias=False))

        def temporal(cin, cout):
            return spectral_norm(nn.Conv3d(cin, cout, kernel_size=(3, 1, 1), stride=(2, 1, 1), padding=(1, 0, 0),
                                           bias=False))

        block = [R2Plus1dStem()]
        for make, cin, cout in ((spatial, 64, 128), (temporal, 128, 128), (spatial, 128, 128), (temporal, 128, 256),
                                (spatial, 256, 256), (temporal, 256, 512), (spatial, 512, 512),
                                (temporal, 512, 512)):
            block += [make(cin, cout), nn.BatchNorm3d(cout), nn.LeakyReLU(0.2)]
        self.pool = nn.AdaptiveAvgPool3d(1)
        self.story_encoder = nn.Sequential(*block)
        self.detector = nn.Sequential(
            spectral_norm(nn.Linear(512, 128)),
            nn.BatchNorm1d(128),
            nn.ReLU(),
            spectral_norm(nn.Linear(128, 1)))

    def forward(self, story):
        from cpcsv_b200 import video
        need_grad = torch.is_grad_enabled() and (
            story.requires_grad or any(p.requires_grad for p in self.parameters()))
        latents = video.VideoEncoderRunner(self, need_grad).apply(story)
        det = self.detector
        x = Fx.linear(latents, Fx.spectral_weight(det[0]), det[0].bias)
        x = torch.relu(Fx.batch_norm_1d(x, det[1]))
        return Fx.linear(x, Fx.spectral_weight(det[3]), det[3].bias)


# --------------------------------------------------------------------------- generator
class StoryGAN(nn.Module):
    """CP-CSV generator with the figure-ground segmentation branch (reference model.py:214-483)."""

    def __init__(self, video_len):
        super(StoryGAN, self).__init__()
        self.batch_size = cfg.TRAIN.IM_BATCH_SIZE
        self.gf_dim = cfg.GAN.GF_DIM * 8
        self.gf_dim_seg = cfg.GAN.GF_SEG_DIM
        self.motion_dim = cfg.TEXT.DIMENSION + cfg.LABEL_NUM
        self.content_dim = cfg.GAN.CONDITION_DIM
        self.noise_dim = cfg.GAN.Z_DIM
        self.recurrent = nn.GRUCell(self.noise_dim + self.motion_dim, self.motion_dim)
        self.mocornn = nn.GRUCell(self.motion_dim, self.content_dim)
        self.video_len = video_len
        self.n_channels = 3
        self.filter_num = 3
        self.filter_size = 21
        self.image_size = 124
        self.out_num = 1
        self.use_segment = cfg.SEGMENT_LEARNING
        self.segment_size = 64
        self.segment_flat_size = 3 * self.segment_size ** 2
        self.aux_size = 5
        self._cpcsv_maps = {}
        self.define_module()

    def define_module(self):
        from layers import DynamicFilterLayer1D
        ninput = self.motion_dim + self.content_dim + self.image_size
        ngf = self.gf_dim
        self.ca_net = CA_NET()
        self.filter_net = _linear_bn(self.content_dim, self.filter_size * self.filter_num * self.out_num)
        self.image_net = _linear_bn(self.motion_dim, self.image_size * self.filter_num, tail=nn.Tanh())
        self.fc = _linear_bn(ninput, ngf * 16, bias=False, tail=nn.ReLU(True))
        self.upsample1 = upBlock(ngf, ngf // 2)
        self.upsample2 = upBlock(ngf // 2, ngf // 4)
        self.upsample3 = upBlock(ngf // 4, ngf // 8)
        self.upsample4 = upBlock(ngf // 8, ngf // 16)
        self.img = nn.Sequential(conv3x3(ngf // 16, 3), nn.Tanh())
        if not self.use_segment:
            raise NotImplementedError("cfg.SEGMENT_LEARNING=False is not part of the accelerated path "
                                      "(cfg/final.yml:19 enables it)")
        nseg = self.gf_dim_seg
        self.seg_c = conv3x3(nseg, ngf)
        self.seg_c1 = conv3x3(nseg // 2, ngf // 2)
        self.fc_seg = _linear_bn(ninput, nseg * 16, bias=False, tail=nn.ReLU(True))
        self.upsample1_seg = upBlock(nseg, nseg // 2)
        self.upsample2_seg = upBlock(nseg // 2, nseg // 4)
        self.upsample3_seg = upBlock(nseg // 4, nseg // 8)
        self.upsample4_seg = upBlock(nseg // 8, nseg // 16)
        self.img_seg = nn.Sequential(conv3x3(nseg // 16, 1), nn.Tanh())
        self.m_net = _linear_bn(self.motion_dim, self.motion_dim)
        self.c_net = _linear_bn(self.content_dim, self.content_dim)
        self.dfn_layer = DynamicFilterLayer1D(self.filter_size, pad=self.filter_size // 2)

    # ---- noise sources (public in the reference; parity tests override them) -------------
    def get_iteration_input(self, motion_input):
        noise = torch.randn(motion_input.shape[0], self.noise_dim, device=motion_input.device)
        return torch.cat((noise, motion_input), dim=1)

    def get_gru_initial_state(self, num_samples):
        dev = self.recurrent.weight_ih.device
        return torch.randn(num_samples, self.motion_dim, device=dev)

    # ---- conditioning path (fp32 kernels) ------------------------------------------------
    def _lin_bn(self, seq, x, tanh=False):
        lin, bn = seq[0], seq[1]
        return Fx.batch_norm_1d(Fx.linear(x, lin.weight, lin.bias), bn, act_tanh=tanh)

    def sample_z_motion(self, motion_input, video_len=None):
        video_len = video_len if video_len is not None else self.video_len
        h0 = self._lin_bn(self.m_net, self.get_gru_initial_state(motion_input.shape[0]))
        # the step inputs (fresh noise | motion_t) do not depend on the recurrence: draw them in
        # the reference's order, then run the whole sequence as one op
        xs = [self.get_iteration_input(motion_input if motion_input.dim() == 2 else motion_input[:, t, :])
              for t in range(video_len)]
        h_all = Fx.gru_sequence(torch.stack(xs, 0), h0, self.recurrent)            # [T, B, H]
        return h_all.transpose(0, 1).reshape(-1, self.motion_dim)

    def motion_content_rnn(self, motion_input, content_input):
        h0 = self._lin_bn(self.c_net, content_input)
        if motion_input.dim() == 2:
            motion_input = motion_input.unsqueeze(1)
            steps = 1
        else:
            steps = self.video_len
        h_all = Fx.gru_sequence(motion_input[:, :steps].transpose(0, 1), h0, self.mocornn)   # [T, B, H]
        return h_all.transpose(0, 1).reshape(-1, self.content_dim)

    def _conditioning(self, motion_input, motion_flat, content_code, c_mu_rows, video_len):
        """everything between the text embeddings and the trunk input (reference model.py:363-378 /
        436-443) as three independent chains on parallel streams: context GRU -> filter_net,
        motion GRU, image_net; then the dynamic filter and the concatenation.  The chains are
        issued in the reference's order, so noise is drawn in the reference's order."""
        crnn_filter, zm_code, m_image = streams.concurrently(
            lambda: self._lin_bn(self.filter_net, self.motion_content_rnn(motion_input, content_code)),
            lambda: self.sample_z_motion(motion_input, video_len),
            lambda: self._lin_bn(self.image_net, motion_flat, tanh=True))
        m_image = m_image.reshape(-1, self.filter_num, self.image_size)
        c_filter = crnn_filter.reshape(-1, self.out_num, self.filter_num, self.filter_size)
        mc_image = self.dfn_layer([m_image, c_filter])
        return torch.cat((zm_code, c_mu_rows, mc_image.squeeze(1)), dim=1)

    def _trunk(self, zmc_all, seg):
        need_grad = torch.is_grad_enabled() and (
            zmc_all.requires_grad or any(p.requires_grad for p in self.parameters()))
        img, segm = nets.TrunkRunner(self, need_grad, seg).apply(zmc_all)
        return None, img, segm      # no latents (the cascade variant returns them, cascade_model.py)

    # ---- public sampling API -------------------------------------------------------------
    def sample_videos(self, motion_input, content_input, seg=False):
        """motion_input (B, V, text+label), content_input (B, V, text) -> 7-tuple
        (None, fake (B,3,V,64,64), m_mu, m_logvar, r_mu, r_logvar, seg-or-None).  Reproduces the
        reference's ``r_mu.repeat(V, 1)`` row order (model.py:361)."""
        B, V = motion_input.shape[0], motion_input.shape[1]
        content = content_input.reshape(B, cfg.VIDEO_LEN * content_input.shape[2])
        r_code, r_mu, r_logvar = self.ca_net(content)
        c_mu = r_mu.repeat(self.video_len, 1)
        m_flat = motion_input.reshape(-1, motion_input.shape[2])
        zmc_all = self._conditioning(motion_input, m_flat, r_code, c_mu, self.video_len)
        latents, img, segm = self._trunk(zmc_all, seg)
        fake = img.view(B, self.video_len, self.n_channels, self.segment_size, self.segment_size)
        fake = fake.permute(0, 2, 1, 3, 4)
        return latents, fake, m_flat, m_flat, r_mu, r_logvar, (segm if seg else None)

    def sample_images(self, motion_input, content_input, seg=False):
        """motion_input (N, text+label), content_input (N, V, text).  The context GRU is seeded
        with c_mu, not the sampled code (reference model.py:433)."""
        N = motion_input.shape[0]
        content = content_input.reshape(N, cfg.VIDEO_LEN * content_input.shape[2])
        _c_code, c_mu, c_logvar = self.ca_net(content)
        zmc_all = self._conditioning(motion_input, motion_input, c_mu, c_mu, 1)
        latents, img, segm = self._trunk(zmc_all, seg)
        return latents, img, motion_input, motion_input, c_mu, c_logvar, (segm if seg else None)


# --------------------------------------------------------------------------- discriminators
def _encoder(in_ch, ndf, sn_first):
    first = nn.Conv2d(in_ch, ndf, 4, 2, 1, bias=False)
    layers = [spectral_norm(first) if sn_first else first, nn.LeakyReLU(0.2, inplace=True)]
    c = ndf
    for _ in range(3):
        layers += [spectral_norm(nn.Conv2d(c, c * 2, 4, 2, 1, bias=False)), nn.BatchNorm2d(c * 2),
                   nn.LeakyReLU(0.2, inplace=True)]
        c *= 2
    return nn.Sequential(*layers)


class _DiscriminatorBase(nn.Module):
    in_channels = 3
    sn_first = False
    categories = True

    def __init__(self, use_categories=True):
        super(_DiscriminatorBase, self).__init__()
        self.df_dim = cfg.GAN.DF_DIM
        self.ef_dim = cfg.GAN.CONDITION_DIM
        self.text_dim = cfg.TEXT.DIMENSION
        self.label_num = cfg.LABEL_NUM
        self.define_module(use_categories and self.categories)

    def define_module(self, use_categories):
        ndf, nef = self.df_dim, self.ef_dim
        self.encode_img = _encoder(self.in_channels, ndf, self.sn_first)
        self.seq_consisten_model = None
        self.get_cond_logits = D_GET_LOGITS(ndf, nef + self.text_dim + self.label_num)
        self.get_uncond_logits = None
        self.cate_classify = CateClassifyConv2d(ndf * 8, self.label_num, 4, 4, 1, bias=False) \
            if use_categories else None

    def _encode(self, image):
        need_grad = torch.is_grad_enabled() and (
            image.requires_grad or any(p.requires_grad for p in self.encode_img.parameters()))
        return nets.EncoderRunner(self, need_grad).apply(image)

    def forward(self, image):
        return self._encode(image)


class STAGE1_D_IMG(_DiscriminatorBase):
    """Image discriminator (reference model.py:487-527)."""


class STAGE1_D_SEG(_DiscriminatorBase):
    """Segmentation-mask discriminator, 1 input channel (reference model.py:529-569)."""
    in_channels = 1


class STAGE1_D_STY_V2(_DiscriminatorBase):
    """Story discriminator: per-frame encoder (spectral norm also on layer 0), mean over the
    V frames (reference model.py:571-618)."""
    sn_first = True
    categories = False

    def __init__(self):
        super(STAGE1_D_STY_V2, self).__init__(use_categories=False)
        if cfg.USE_SEQ_CONSISTENCY:       # reference model.py:599-601
            self.seq_consisten_model = VideoEncoder()

    def forward(self, story):
        N, C, video_len, W, H = story.shape
        frames = story.permute(0, 2, 1, 3, 4).contiguous().view(-1, C, W, H)
        emb = torch.squeeze(self._encode(frames))
        _, C1, W1, H1 = emb.shape
        return emb.view(N, video_len, C1, W1, H1).mean(1).squeeze()
