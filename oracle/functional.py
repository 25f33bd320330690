"""fp32 functional restatement of the CP-CSV hot path (TEST INFRASTRUCTURE -- see
``oracle/__init__.py``).

Every network is a plain ``dict`` name -> tensor with the reference's ``state_dict`` keys
(SURVEY.md Appendix B); functions below take that dict and compute with
``torch.nn.functional`` calls in fp32.  Citations are into ``/root/reference``.

Stateful side effects the reference has are reproduced in place on the dict's tensors:
BatchNorm running statistics (every forward, also under no_grad) and the spectral-norm
power iteration (one per forward call of the wrapped conv, ``u``/``v`` persisted).
"""
import torch
import torch.nn.functional as F

IMAGE_SIZE = 124   # model.py:229
FILTER_NUM = 3     # model.py:227
FILTER_SIZE = 21   # model.py:228


# --------------------------------------------------------------------------- primitives
_TRAINING = [True]


class eval_mode:
    """``with eval_mode():`` -- what ``netG.eval()`` changes (reference inference.py:88,
    trainer.py:161,177): every BatchNorm normalises with its running statistics and leaves them
    untouched.  (The generator has no other mode-dependent layer.)"""

    def __enter__(self):
        self.prev = _TRAINING[0]
        _TRAINING[0] = False

    def __exit__(self, *exc):
        _TRAINING[0] = self.prev
        return False


def batch_norm(sd, prefix, x):
    """nn.BatchNorm1d/2d in train mode, eps 1e-5, momentum 0.1 (model.py:32,252,...)."""
    if not _TRAINING[0]:
        return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                            sd[prefix + ".weight"], sd[prefix + ".bias"], training=False, eps=1e-5)
    sd[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"],
                        training=True, momentum=0.1, eps=1e-5)


def spectral_weight(sd, prefix):
    """Legacy ``torch.nn.utils.spectral_norm`` hook, train mode, 1 power iteration,
    eps 1e-12 (model.py:5,19,79; torch/nn/utils/spectral_norm.py compute_weight)."""
    w = sd[prefix + ".weight_orig"]
    u, v = sd[prefix + ".weight_u"], sd[prefix + ".weight_v"]
    wm = w.reshape(w.shape[0], -1)
    with torch.no_grad():
        v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12))
        u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=1e-12))
        uc, vc = u.clone(), v.clone()
    sigma = torch.dot(uc, torch.mv(wm, vc))
    return w / sigma


def up_block(sd, name, x):
    """upBlock: nearest x2 -> conv3x3 (no bias) -> BN -> ReLU (model.py:26-34)."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = F.conv2d(x, sd[name + ".1.weight"], None, 1, 1)
    return F.relu(batch_norm(sd, name + ".2", x))


def gru_cell(sd, name, x, h):
    """nn.GRUCell, gate order (r, z, n) (model.py:223-224)."""
    return torch.gru_cell(x, h, sd[name + ".weight_ih"], sd[name + ".weight_hh"],
                          sd[name + ".bias_ih"], sd[name + ".bias_hh"])


def dynamic_filter_1d(image, filters):
    """DynamicFilterLayer1D (layers.py:62-80): per-sample cross-correlation, zero pad K//2.
    image (N,3,L), filters (N,1,3,K) -> (N,1,L)."""
    n, c, length = image.shape
    k = filters.shape[-1]
    out = F.conv1d(image.reshape(1, n * c, length), filters.reshape(n, c, k),
                   padding=k // 2, groups=n)
    return out.reshape(n, 1, length)


# --------------------------------------------------------------------------- generator
def ca_net(sd, text, eps):
    """CA_NET.forward (model.py:47-65); ``eps`` is the injected N(0,1) draw."""
    x = F.relu(F.linear(text, sd["ca_net.fc.weight"], sd["ca_net.fc.bias"]))
    c = x.shape[1] // 2
    mu, logvar = x[:, :c], x[:, c:]
    std = torch.exp(0.5 * logvar)
    return eps * std + mu, mu, logvar


def linear_bn(sd, name, x):
    y = F.linear(x, sd[name + ".0.weight"], sd.get(name + ".0.bias"))
    return batch_norm(sd, name + ".1", y)


def motion_content_rnn(sd, motion, content):
    """model.py:336-346."""
    h = linear_bn(sd, "c_net", content)
    if motion.dim() == 2:
        motion = motion.unsqueeze(1)
    outs = []
    for t in range(motion.shape[1]):
        h = gru_cell(sd, "mocornn", motion[:, t], h)
        outs.append(h)
    return torch.stack(outs, 1).reshape(-1, h.shape[1])


def sample_z_motion(sd, motion, video_len, noise):
    """model.py:313-334; draws h0 noise first, then one noise tensor per frame."""
    n = motion.shape[0]
    m_dim = sd["m_net.0.weight"].shape[0]
    z_dim = sd["recurrent.weight_ih"].shape[1] - m_dim
    h = linear_bn(sd, "m_net", noise.pop((n, m_dim)))
    outs = []
    for t in range(video_len):
        m_t = motion if motion.dim() == 2 else motion[:, t]
        e_t = torch.cat((noise.pop((n, z_dim)), m_t), 1)
        h = gru_cell(sd, "recurrent", e_t, h)
        outs.append(h)
    return torch.stack(outs, 1).reshape(-1, m_dim)


def _trunk(sd, zmc_all, use_segment=True):
    """fc/fc_seg + the two up-sampling trunks + heads (model.py:379-407 / 445-470)."""
    ngf = sd["upsample1.1.weight"].shape[1]
    zmc_img = F.relu(batch_norm(sd, "fc.1", F.linear(zmc_all, sd["fc.0.weight"]))).view(-1, ngf, 4, 4)
    if not use_segment:
        h = zmc_img
        for i in range(1, 5):
            h = up_block(sd, "upsample%d" % i, h)
        return torch.tanh(F.conv2d(h, sd["img.0.weight"], None, 1, 1)), None
    nseg = sd["upsample1_seg.1.weight"].shape[1]
    zmc_seg = F.relu(batch_norm(sd, "fc_seg.1", F.linear(zmc_all, sd["fc_seg.0.weight"]))).view(-1, nseg, 4, 4)
    zmc_img = F.conv2d(zmc_seg, sd["seg_c.weight"], None, 1, 1) * zmc_img + zmc_img
    h_seg = up_block(sd, "upsample1_seg", zmc_seg)
    h_img = up_block(sd, "upsample1", zmc_img)
    h_img = F.conv2d(h_seg, sd["seg_c1.weight"], None, 1, 1) * h_img + h_img
    for i in (2, 3, 4):
        h_seg = up_block(sd, "upsample%d_seg" % i, h_seg)
        h_img = up_block(sd, "upsample%d" % i, h_img)
    seg = torch.tanh(F.conv2d(h_seg, sd["img_seg.0.weight"], None, 1, 1))
    img = torch.tanh(F.conv2d(h_img, sd["img.0.weight"], None, 1, 1))
    return img, seg


def down_block(sd, name, x):
    """downBlock: conv3x3 stride 2 pad 1 WITH bias -> BN -> ReLU (cascade_model.py:36-41)."""
    x = F.conv2d(x, sd[name + ".0.weight"], sd[name + ".0.bias"], 2, 1)
    return F.relu(batch_norm(sd, name + ".1", x))


def seg_encoder(sd, seg):
    """presample + downsample1..4_seg (cascade_model.py:312-320, 413-418): mask (n,1,64,64) ->
    (g_seg1 (4x4), g_seg2 (8x8), g_seg3 (16x16), g_seg4 (32x32))."""
    z = F.conv2d(seg, sd["presample.0.weight"], None, 1, 1)
    z = F.relu(batch_norm(sd, "presample.1", z))
    g4 = down_block(sd, "downsample1_seg", z)
    g3 = down_block(sd, "downsample2_seg", g4)
    g2 = down_block(sd, "downsample3_seg", g3)
    g1 = down_block(sd, "downsample4_seg", g2)
    return g1, g2, g3, g4


def _trunk_cascade(sd, zmc_all):
    """cascade_model.py:399-437 / 477-511: the segmentation trunk runs first, its mask is
    re-encoded by presample + 4 downBlocks, and the image trunk is modulated by the
    re-encoded features g_seg1 / g_seg2.  Returns (img, seg, latents) with latents =
    ((zmc_seg, h_seg1, h_seg2, h_seg3), (g_seg1, g_seg2, g_seg3, g_seg4))."""
    ngf = sd["upsample1.1.weight"].shape[1]
    nseg = sd["upsample1_seg.1.weight"].shape[1]
    zmc_img = F.relu(batch_norm(sd, "fc.1", F.linear(zmc_all, sd["fc.0.weight"]))).view(-1, ngf, 4, 4)
    zmc_seg = F.relu(batch_norm(sd, "fc_seg.1", F.linear(zmc_all, sd["fc_seg.0.weight"]))).view(-1, nseg, 4, 4)
    h1 = up_block(sd, "upsample1_seg", zmc_seg)
    h2 = up_block(sd, "upsample2_seg", h1)
    h3 = up_block(sd, "upsample3_seg", h2)
    h4 = up_block(sd, "upsample4_seg", h3)
    seg = torch.tanh(F.conv2d(h4, sd["img_seg.0.weight"], None, 1, 1))
    g1, g2, g3, g4 = seg_encoder(sd, seg)
    zmc_img = F.conv2d(g1, sd["seg_c.weight"], None, 1, 1) * zmc_img + zmc_img
    h_img = up_block(sd, "upsample1", zmc_img)
    h_img = F.conv2d(g2, sd["seg_c1.weight"], None, 1, 1) * h_img + h_img
    for i in (2, 3, 4):
        h_img = up_block(sd, "upsample%d" % i, h_img)
    img = torch.tanh(F.conv2d(h_img, sd["img.0.weight"], None, 1, 1))
    return img, seg, ((zmc_seg, h1, h2, h3), (g1, g2, g3, g4))


def train_autoencoder(sd, real_segments):
    """StoryGAN.train_autoencoder (cascade_model.py:528-540): mask -> encoder -> seg up-blocks
    -> img_seg."""
    g1, _g2, _g3, _g4 = seg_encoder(sd, real_segments)
    h = g1
    for i in (1, 2, 3, 4):
        h = up_block(sd, "upsample%d_seg" % i, h)
    return torch.tanh(F.conv2d(h, sd["img_seg.0.weight"], None, 1, 1))


def is_cascade(sd):
    return "presample.0.weight" in sd


def _run_trunk(sd, zmc_all, use_segment):
    if is_cascade(sd):
        return _trunk_cascade(sd, zmc_all)
    img, seg = _trunk(sd, zmc_all, use_segment)
    return img, seg, None


def _cond_to_latent(sd, motion_flat, crnn_code, zm_code, c_mu_rows):
    """model.py:371-378 / 436-443: image_net, filter_net, dynamic filter, concat."""
    zmc_code = torch.cat((zm_code, c_mu_rows), 1)
    m_image = torch.tanh(linear_bn(sd, "image_net", motion_flat)).view(-1, FILTER_NUM, IMAGE_SIZE)
    c_filter = linear_bn(sd, "filter_net", crnn_code).view(-1, 1, FILTER_NUM, FILTER_SIZE)
    mc_image = dynamic_filter_1d(m_image, c_filter)
    return torch.cat((zmc_code, mc_image.squeeze(1)), 1)


def sample_videos(sd, motion_input, content_input, noise, seg=False, use_segment=True):
    """StoryGAN.sample_videos (model.py:348-423).  motion (B,V,M), content (B,V,T).
    Note the reference's row order quirk at model.py:361: ``r_mu.repeat(V,1)`` puts
    r_mu[j mod B] in row j although every other per-frame tensor is ordered b*V+t."""
    B, V = motion_input.shape[0], motion_input.shape[1]
    content = content_input.reshape(B, -1)
    c_dim = sd["c_net.0.weight"].shape[0]
    r_code, r_mu, r_logvar = ca_net(sd, content, noise.pop((B, c_dim)))
    c_mu = r_mu.repeat(V, 1)
    crnn_code = motion_content_rnn(sd, motion_input, r_code)
    m_flat = motion_input.reshape(B * V, -1)
    zm_code = sample_z_motion(sd, motion_input, V, noise)
    zmc_all = _cond_to_latent(sd, m_flat, crnn_code, zm_code, c_mu)
    img, segm, latents = _run_trunk(sd, zmc_all, use_segment)
    fake = img.view(B, V, 3, 64, 64).permute(0, 2, 1, 3, 4)
    return latents, fake, m_flat, m_flat, r_mu, r_logvar, (segm if seg else None)


def sample_images(sd, motion_input, content_input, noise, seg=False, use_segment=True):
    """StoryGAN.sample_images (model.py:426-483).  motion (N,M), content (N,V,T).  The
    context GRU is seeded with c_mu, not c_code (model.py:433)."""
    N = motion_input.shape[0]
    content = content_input.reshape(N, -1)
    c_dim = sd["c_net.0.weight"].shape[0]
    _c_code, c_mu, c_logvar = ca_net(sd, content, noise.pop((N, c_dim)))
    crnn_code = motion_content_rnn(sd, motion_input, c_mu)
    zm_code = sample_z_motion(sd, motion_input, 1, noise)
    zmc_all = _cond_to_latent(sd, motion_input, crnn_code, zm_code, c_mu)
    img, segm, latents = _run_trunk(sd, zmc_all, use_segment)
    return latents, img, motion_input, motion_input, c_mu, c_logvar, (segm if seg else None)


# ----------------------------------------------------------------------- discriminators
def encode_img(sd, x):
    """4x [conv4x4 s2 p1 (no bias) [SN] -> [BN] -> LeakyReLU(0.2)] (model.py:498-514,
    540-556, 582-598).  Layer 0 has SN only in D_STY and never BN."""
    if "encode_img.0.weight_orig" in sd:
        w0 = spectral_weight(sd, "encode_img.0")
    else:
        w0 = sd["encode_img.0.weight"]
    h = F.leaky_relu(F.conv2d(x, w0, None, 2, 1), 0.2)
    for idx in (2, 5, 8):
        h = F.conv2d(h, spectral_weight(sd, "encode_img.%d" % idx), None, 2, 1)
        h = F.leaky_relu(batch_norm(sd, "encode_img.%d" % (idx + 1), h), 0.2)
    return h


def d_forward(sd, x):
    """STAGE1_D_IMG/SEG.forward (model.py:524-527) for 4-D input; STAGE1_D_STY_V2.forward
    (model.py:610-618) for 5-D (N,C,V,H,W) stories: per-frame encoder then mean over V."""
    if x.dim() == 4:
        return encode_img(sd, x)
    n, c, v, hgt, wid = x.shape
    frames = x.permute(0, 2, 1, 3, 4).contiguous().view(-1, c, hgt, wid)
    emb = torch.squeeze(encode_img(sd, frames))
    return emb.view(n, v, *emb.shape[1:]).mean(1).squeeze()


def get_cond_logits(sd, h_code, c_code):
    """D_GET_LOGITS.forward with bcondition (model.py:86-97)."""
    p = "get_cond_logits.outlogits"
    c = c_code.view(c_code.shape[0], -1, 1, 1).repeat(1, 1, 4, 4)
    x = torch.cat((h_code, c), 1)
    x = F.conv2d(x, spectral_weight(sd, p + ".0"), None, 1, 1)
    x = F.leaky_relu(batch_norm(sd, p + ".1", x), 0.2)
    x = F.conv2d(x, spectral_weight(sd, p + ".3"), sd[p + ".3.bias"], 4, 0)
    return torch.sigmoid(x).view(-1)


def cate_classify(sd, h_code):
    """nn.Conv2d(ndf*8, label_num, 4, 4, 1, bias=False) (model.py:520)."""
    return F.conv2d(h_code, sd["cate_classify.weight"], None, 4, 1)


# ------------------------------------------------------------------------------ losses
def multi_acc(logits, labels):
    """get_multi_acc (miscc/utils.py:313-321)."""
    hit = ((torch.sigmoid(logits) >= 0.5) & (labels == 1)).sum()
    return (hit.float() / labels.sum()).item()


def discriminator_loss(sd, real_imgs, fake_imgs, real_labels, fake_labels, cate_labels, cond, consistency_ratio=1.0):
    """compute_discriminator_loss, conditional branch (miscc/utils.py:48-123).  Returns
    (errD, errD_real, errD_wrong, errD_fake, cate_logits-or-None)."""
    n = real_imgs.size(0)
    cond = cond.detach()
    real_f = d_forward(sd, real_imgs)
    fake_f = d_forward(sd, fake_imgs.detach())
    e_real = F.binary_cross_entropy(get_cond_logits(sd, real_f, cond), real_labels)
    e_wrong = F.binary_cross_entropy(get_cond_logits(sd, real_f[:n - 1], cond[1:]), fake_labels[1:])
    e_fake = F.binary_cross_entropy(get_cond_logits(sd, fake_f, cond), fake_labels)
    err = e_real + (e_fake + e_wrong) * 0.5
    cate = None
    if "cate_classify.weight" in sd:
        cate = cate_classify(sd, real_f).squeeze()
        err = err + 1.0 * F.multilabel_soft_margin_loss(cate, cate_labels)
    if "seq_consisten_model.detector.3.weight_orig" in sd:
        # miscc/utils.py:110-122 (cfg.USE_SEQ_CONSISTENCY)
        from . import video_encoder as VE
        shuffled, order_labels = VE.create_random_shuffle(real_imgs)
        order, _ = VE.order_loss_d(VE.sub_state(sd), shuffled, order_labels)
        err = err + consistency_ratio * order
    return err, e_real.detach(), e_wrong.detach(), e_fake.detach(), cate


def generator_loss(sd, fake_imgs, real_labels, cate_labels, cond, real_imgs=None, consistency_ratio=1.0):
    """compute_generator_loss, conditional branch (miscc/utils.py:126-171)."""
    cond = cond.detach()
    fake_f = d_forward(sd, fake_imgs)
    err = F.binary_cross_entropy(get_cond_logits(sd, fake_f, cond), real_labels)
    cate = None
    if "cate_classify.weight" in sd:
        cate = cate_classify(sd, fake_f).squeeze()
        err = err + 1.0 * F.multilabel_soft_margin_loss(cate, cate_labels)
    if "seq_consisten_model.detector.3.weight_orig" in sd:
        # miscc/utils.py:155-169
        from . import video_encoder as VE
        err = err + consistency_ratio * VE.order_loss_g(VE.sub_state(sd), real_imgs, fake_imgs)
    return err, cate


def kl_loss(mu, logvar):
    """KL_loss (miscc/utils.py:184-188)."""
    return -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())


# -------------------------------------------------------------------------------- step
class OracleModel:
    """The four state dicts as leaf tensors + the four Adam optimisers of
    trainer.py:212-220 (betas (0.5, 0.999); lr D 4e-4, G 1e-4 at cfg/final.yml)."""

    def __init__(self, states, p, device="cpu", with_optim=True):
        from .params import is_parameter
        self.p = p
        self.nets = {}
        for net, sd in states.items():
            out = {}
            for k, v in sd.items():
                t = v.detach().clone().to(device)
                if is_parameter(k):
                    t.requires_grad_(True)
                out[k] = t
            self.nets[net] = out
        self.opt = {}
        if with_optim:
            for net in self.nets:
                lr = p["GENERATOR_LR"] if net == "G" else p["DISCRIMINATOR_LR"]
                self.opt[net] = torch.optim.Adam(self.params(net), lr=lr, betas=(0.5, 0.999))

    def params(self, net):
        return [t for t in self.nets[net].values() if t.requires_grad]

    def named_params(self, net):
        return [(k, t) for k, t in self.nets[net].items() if t.requires_grad]

    def zero_grad(self, net):
        for t in self.params(net):
            t.grad = None


def train_step(model, batch, noise, ratio=1.0, apply_optim=True):
    """One iteration of GANTrainer.train (trainer.py:252-416), SEGMENT_LEARNING on; with a
    cascade generator state dict (CASCADE_MODEL, cascade_model.py) also the latent-MSE and
    reconstruction losses of trainer.py:369-384, 412-413.  ``noise`` is a synth.NoiseFeed.  Returns a dict with the losses,
    the generated tensors of both generator passes and (via ``.grad``) all gradients."""
    p = model.p
    G, D_im, D_st, D_se = (model.nets[k] for k in ("G", "D_im", "D_st", "D_se"))
    T = p["TEXT_DIM"]
    st_real, im_real, se_real = batch["st_real"], batch["im_real"], batch["se_real"]
    st_labels, im_labels = batch["st_labels"], batch["im_labels"]
    # trainer.py:255-288
    im_motion = torch.cat((batch["im_desc"][:, :T], im_labels), 1)
    im_content = batch["im_content"][:, :, :T]
    st_motion = torch.cat((batch["st_desc"][:, :, :T], st_labels), 2)
    st_content = batch["st_desc"][:, :, :T]
    dev = st_real.device
    dt = st_real.dtype
    im_ones = torch.ones(im_real.shape[0], device=dev, dtype=dt)
    im_zeros = torch.zeros(im_real.shape[0], device=dev, dtype=dt)
    st_ones = torch.ones(st_real.shape[0], device=dev, dtype=dt)
    st_zeros = torch.zeros(st_real.shape[0], device=dev, dtype=dt)
    out = {}

    def cond_vectors(c_mu, cim_mu):
        # trainer.py:303-307 / 386-389
        chars = (st_labels.mean(1) > 0).to(st_labels.dtype)
        st_mu = torch.cat((c_mu, st_motion[:, :, :T].mean(1).squeeze(), chars), 1)
        im_mu = torch.cat((im_motion, cim_mu), 1)
        return st_mu, im_mu

    # (2) no-grad generator passes, trainer.py:295-300
    with torch.no_grad():
        _, st_fake, _, _, c_mu, _c_lv, _ = sample_videos(G, st_motion, st_content, noise)
        _, im_fake, _, _, cim_mu, _cim_lv, se_fake = sample_images(G, im_motion, im_content, noise, seg=True)
    st_mu, im_mu = cond_vectors(c_mu, cim_mu)
    out["p1_st_fake"], out["p1_im_fake"], out["p1_se_fake"] = st_fake, im_fake, se_fake

    # (3) discriminator update, trainer.py:313-346
    for net in ("D_im", "D_st", "D_se"):
        model.zero_grad(net)
    se_errD, *_ = discriminator_loss(D_se, se_real, se_fake, im_ones, im_zeros, im_labels, im_mu)
    im_errD, *_ = discriminator_loss(D_im, im_real, im_fake, im_ones, im_zeros, im_labels, im_mu)
    st_errD, *_ = discriminator_loss(D_st, st_real, st_fake, st_ones, st_zeros, st_labels, st_mu,
                                     p.get("CONSISTENCY_RATIO", 1.0))
    se_errD.backward()
    if apply_optim:
        model.opt["D_se"].step()
    im_errD.backward()
    st_errD.backward()
    if apply_optim:
        model.opt["D_im"].step()
        model.opt["D_st"].step()
    out.update(se_errD=se_errD.detach(), im_errD=im_errD.detach(), st_errD=st_errD.detach())
    out["D_grads"] = {net: {k: t.grad.detach().clone() for k, t in model.named_params(net)}
                      for net in ("D_im", "D_st", "D_se")}

    # (4) generator update, trainer.py:365-416
    model.zero_grad("G")
    video_latents, st_fake, _, _, c_mu, c_logvar, _ = sample_videos(G, st_motion, st_content, noise)
    image_latents, im_fake, _, _, cim_mu, cim_logvar, se_fake = sample_images(G, im_motion, im_content, noise,
                                                                              seg=True)
    if video_latents is not None:
        # trainer.py:370-380.  The reference unpacks the first tuple as (h_seg1..4) although it
        # holds (zmc_seg, h_seg1, h_seg2, h_seg3): pairs are matched by position.
        def latent_loss(latents):
            hs, gs = latents
            return sum(F.mse_loss(g, h) for g, h in zip(gs, hs))
        video_latent_loss = latent_loss(video_latents)
        image_latent_loss = latent_loss(image_latents)
        rec_real = train_autoencoder(G, se_real)
        rec_fake = train_autoencoder(G, se_fake)
        reconstruct_loss = (F.mse_loss(rec_real, se_real) + F.mse_loss(rec_fake, se_fake)) / 2.0
        out.update(video_latent_loss=video_latent_loss.detach(), image_latent_loss=image_latent_loss.detach(),
                   reconstruct_loss=reconstruct_loss.detach())
    st_mu, im_mu = cond_vectors(c_mu, cim_mu)
    se_errG, _ = generator_loss(D_se, se_fake, im_ones, im_labels, im_mu)
    im_errG, _ = generator_loss(D_im, im_fake, im_ones, im_labels, im_mu)
    st_errG, _ = generator_loss(D_st, st_fake, st_ones, st_labels, st_mu, st_real, p.get("CONSISTENCY_RATIO", 1.0))
    im_kl = kl_loss(cim_mu, cim_logvar)
    st_kl = kl_loss(c_mu, c_logvar)
    kl_w = p["KL"]
    total = im_errG + im_kl * kl_w + ratio * (se_errG * p["SEGMENT_RATIO"] + st_errG * p["IMAGE_RATIO"]
                                                + st_kl * kl_w)
    if video_latents is not None:
        total = total + (video_latent_loss + reconstruct_loss) * p.get("RECONSTRUCT_LOSS", 1.0)   # trainer.py:412-413
    total.backward()
    if apply_optim:
        model.opt["G"].step()
    out.update(se_errG=se_errG.detach(), im_errG=im_errG.detach(), st_errG=st_errG.detach(),
               im_kl=im_kl.detach(), st_kl=st_kl.detach(), errG_total=total.detach())
    out["p3_st_fake"], out["p3_im_fake"], out["p3_se_fake"] = (
        st_fake.detach(), im_fake.detach(), se_fake.detach())
    out["G_grads"] = {k: t.grad.detach().clone() for k, t in model.named_params("G")}
    return out
