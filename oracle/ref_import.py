"""Import the REAL reference (``/root/reference``) in the build container.

Only ``oracle/make_golden.py`` uses this, in its own process: the reference's top-level
module names (``model``, ``layers``, ``miscc``) collide with the product's drop-in modules,
so the two are never imported into one interpreter.  ``/root/reference`` does not exist on
the GPU box; nothing under ``tests/ -m gpu``, ``smoke()`` or ``bench.py`` imports this file.

Shims: a stand-in ``easydict``; identity ``Tensor.cuda`` on CPU-only
hosts (``model.py:238`` calls ``.cuda()`` unconditionally); ``data_parallel`` pass-through
(its 1-GPU behaviour).  ``trainer.py`` itself cannot be imported on torch 2.11, so the step
sequence of ``trainer.py:252-416`` is restated in ``reference_step`` around the reference's
own model and loss functions.
"""
import sys
import types
import torch
import torch.nn as nn

REF = "/root/reference"


def _install_shims():
    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                self[k] = v

        def __setitem__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)

        __setattr__ = __setitem__

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

    mod = types.ModuleType("easydict")
    mod.EasyDict = EasyDict
    sys.modules.setdefault("easydict", mod)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    nn.parallel.data_parallel = lambda m, inp, gpus=None: m(*inp) if isinstance(inp, tuple) else m(inp)


def load_reference(p):
    """Returns (model_module, utils_module, cfg) with cfg set from preset ``p``."""
    from . import presets
    _install_shims()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from miscc.config import cfg
    presets.apply_to_cfg(cfg, p)
    cfg.CUDA = torch.cuda.is_available()
    if p.get("CASCADE_MODEL"):
        import cascade_model as ref_model       # trainer.py:83-84
    else:
        import model as ref_model
    import miscc.utils as ref_utils
    assert ref_model.__file__.startswith(REF), ref_model.__file__
    return ref_model, ref_utils, cfg


def build_reference_nets(ref_model, p, states):
    """Construct the reference modules in trainer order (trainer.py:87-97) and load the
    oracle-initialised state dicts with strict key checking."""
    nets = {
        "G": ref_model.StoryGAN(p["VIDEO_LEN"]),
        "D_im": ref_model.STAGE1_D_IMG(),
        "D_st": ref_model.STAGE1_D_STY_V2(),
        "D_se": ref_model.STAGE1_D_SEG(),
    }
    for k, net in nets.items():
        net.load_state_dict(states[k], strict=True)
        net.train()
    return nets


def inject_noise(netG, feed):
    """Make the reference consume pre-generated noise: wrap the three public methods that
    draw it (model.py:53-60, 313-319)."""
    def reparametrize(mu, logvar):
        std = logvar.mul(0.5).exp_()
        return feed.pop(tuple(std.shape)).mul(std).add_(mu)

    def get_iteration_input(motion_input):
        n = motion_input.shape[0]
        return torch.cat((feed.pop((n, netG.noise_dim)), motion_input), dim=1)

    def get_gru_initial_state(num_samples):
        return feed.pop((num_samples, netG.motion_dim))

    netG.ca_net.reparametrize = reparametrize
    netG.get_iteration_input = get_iteration_input
    netG.get_gru_initial_state = get_gru_initial_state


def reference_step(nets, ref_utils, cfg, batch, ratio=1.0, opts=None):
    """trainer.py:252-416 restated around the reference's own functions."""
    netG, netD_im, netD_st, netD_se = nets["G"], nets["D_im"], nets["D_st"], nets["D_se"]
    T = cfg.TEXT.DIMENSION
    gpus = [0]
    st_real, im_real, se_real = batch["st_real"], batch["im_real"], batch["se_real"]
    st_labels, im_labels = batch["st_labels"], batch["im_labels"]
    im_motion = torch.cat((batch["im_desc"][:, :T], im_labels), 1)
    im_content = batch["im_content"][:, :, :T]
    st_motion = torch.cat((batch["st_desc"][:, :, :T], st_labels), 2)
    st_content = batch["st_desc"][:, :, :T]
    im_ones, im_zeros = torch.ones(im_real.shape[0]), torch.zeros(im_real.shape[0])
    st_ones, st_zeros = torch.ones(st_real.shape[0]), torch.zeros(st_real.shape[0])
    out = {}
    with torch.no_grad():
        _, st_fake, m_mu, m_logvar, c_mu, c_logvar, _ = netG.sample_videos(st_motion, st_content)
        _, im_fake, im_mu, im_logvar, cim_mu, cim_logvar, se_fake = netG.sample_images(
            im_motion, im_content, seg=True)
    out["p1_st_fake"], out["p1_im_fake"], out["p1_se_fake"] = st_fake, im_fake, se_fake
    characters_mu = (st_labels.mean(1) > 0).type(torch.FloatTensor)
    st_mu = torch.cat((c_mu, st_motion[:, :, :T].mean(1).squeeze(), characters_mu), 1)
    im_mu = torch.cat((im_motion, cim_mu), 1)
    netD_im.zero_grad(); netD_st.zero_grad(); netD_se.zero_grad()
    se_errD = ref_utils.compute_discriminator_loss(netD_se, se_real, se_fake, im_ones, im_zeros,
                                                   im_labels, im_mu, gpus)[0]
    im_errD = ref_utils.compute_discriminator_loss(netD_im, im_real, im_fake, im_ones, im_zeros,
                                                   im_labels, im_mu, gpus)[0]
    st_errD = ref_utils.compute_discriminator_loss(netD_st, st_real, st_fake, st_ones, st_zeros,
                                                   st_labels, st_mu, gpus)[0]
    se_errD.backward()
    if opts:
        opts["D_se"].step()
    im_errD.backward()
    st_errD.backward()
    if opts:
        opts["D_im"].step()
        opts["D_st"].step()
    out.update(se_errD=se_errD.detach(), im_errD=im_errD.detach(), st_errD=st_errD.detach())
    out["D_grads"] = {k: {n: q.grad.detach().clone() for n, q in nets[k].named_parameters()}
                      for k in ("D_im", "D_st", "D_se")}
    netG.zero_grad()
    video_latents, st_fake, m_mu, m_logvar, c_mu, c_logvar, _ = netG.sample_videos(st_motion, st_content)
    image_latents, im_fake, im_mu, im_logvar, cim_mu, cim_logvar, se_fake = netG.sample_images(
        im_motion, im_content, seg=True)
    if video_latents is not None:        # trainer.py:369-384, verbatim pairing
        mse_loss = nn.MSELoss()
        ((h_seg1, h_seg2, h_seg3, h_seg4), (g_seg1, g_seg2, g_seg3, g_seg4)) = video_latents
        video_latent_loss = mse_loss(g_seg1, h_seg1) + mse_loss(g_seg2, h_seg2) + mse_loss(g_seg3, h_seg3) \
            + mse_loss(g_seg4, h_seg4)
        ((h_seg1, h_seg2, h_seg3, h_seg4), (g_seg1, g_seg2, g_seg3, g_seg4)) = image_latents
        image_latent_loss = mse_loss(g_seg1, h_seg1) + mse_loss(g_seg2, h_seg2) + mse_loss(g_seg3, h_seg3) \
            + mse_loss(g_seg4, h_seg4)
        reconstruct_img = netG.train_autoencoder(se_real)
        reconstruct_fake = netG.train_autoencoder(se_fake)
        reconstruct_loss = (mse_loss(reconstruct_img, se_real) + mse_loss(reconstruct_fake, se_fake)) / 2.0
        out.update(video_latent_loss=video_latent_loss.detach(), image_latent_loss=image_latent_loss.detach(),
                   reconstruct_loss=reconstruct_loss.detach())
    characters_mu = (st_labels.mean(1) > 0).type(torch.FloatTensor)
    st_mu = torch.cat((c_mu, st_motion[:, :, :T].mean(1).squeeze(), characters_mu), 1)
    im_mu = torch.cat((im_motion, cim_mu), 1)
    se_errG = ref_utils.compute_generator_loss(netD_se, se_fake, se_real, im_ones, im_labels, im_mu, gpus)[0]
    im_errG = ref_utils.compute_generator_loss(netD_im, im_fake, im_real, im_ones, im_labels, im_mu, gpus)[0]
    st_errG = ref_utils.compute_generator_loss(netD_st, st_fake, st_real, st_ones, st_labels, st_mu, gpus)[0]
    im_kl = ref_utils.KL_loss(cim_mu, cim_logvar)
    st_kl = ref_utils.KL_loss(c_mu, c_logvar)
    total = im_errG + im_kl * cfg.TRAIN.COEFF.KL + ratio * (
        se_errG * cfg.SEGMENT_RATIO + st_errG * cfg.IMAGE_RATIO + st_kl * cfg.TRAIN.COEFF.KL)
    if video_latents is not None:        # trainer.py:412-413
        total = total + (video_latent_loss + reconstruct_loss) * cfg.RECONSTRUCT_LOSS
    total.backward()
    if opts:
        opts["G"].step()
    out.update(se_errG=se_errG.detach(), im_errG=im_errG.detach(), st_errG=st_errG.detach(),
               im_kl=im_kl.detach(), st_kl=st_kl.detach(), errG_total=total.detach())
    out["p3_st_fake"], out["p3_im_fake"], out["p3_se_fake"] = (
        st_fake.detach(), im_fake.detach(), se_fake.detach())
    out["G_grads"] = {n: q.grad.detach().clone() for n, q in netG.named_parameters()}
    return out
