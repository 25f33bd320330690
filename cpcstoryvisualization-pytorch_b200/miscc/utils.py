"""Loss / bookkeeping helpers with the reference's names and call signatures
(``miscc/utils.py`` of the reference: compute_discriminator_loss l.48-123,
compute_generator_loss l.126-171, KL_loss l.184-188, weights_init l.191-201, get_multi_acc
l.313-321, save_model l.323-338, count_param l.431-435); the consumers of the generator's outputs
(image sheets, PNG / npy writers, sampling loops) live in ``miscc/outputs.py`` and are re-exported here
under the reference's names.

Differences that do not change results: the discriminators are called directly instead of
through ``nn.parallel.data_parallel`` (one process drives one GPU here, where data_parallel
degenerates to a plain call), and the classification accuracy is computed on the device and
only converted to a Python float when ``cfg``-independent flag ``SYNC_ACCURACY`` is left on.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from miscc.config import cfg  # noqa: F401

from miscc.outputs import (check_is_order, compute_cyc_loss_img, compute_cyc_loss_txt,  # noqa: F401,E402
                           create_random_shuffle, images_to_numpy, inference_samples, save_all_img,
                           save_image_results, save_img_results, save_story_results, save_test_samples,
                           save_train_samples)

SYNC_ACCURACY = True   # reference behaviour: .cpu().numpy() round trip per call (host sync)
# Real / fake (/ wrong-pair) passes of ONE discriminator on parallel CUDA streams (CPCSV_PARALLEL_PASSES=0: in
# sequence).  OPEN ISSUE: in graph-replayed steps the losses of the story discriminator -- the
# one whose passes run on the step's own stream -- deviate intermittently by 1-10 % to a few discrete alternative
# values and are exact again in the next step.  How often depends on what runs concurrently (tools/diag_run.sh, 60
# step measurements per setting, 'small' presets): plain generator, parallel passes: 0; plain generator, passes in
# sequence: ~10; cascade generator, parallel passes: ~1; cascade generator, passes in sequence: 0.  The race itself has
# not been found; each generator gets the setting that measured clean.
PARALLEL_PASSES = os.environ.get("CPCSV_PARALLEL_PASSES", "1") != "0"


import contextlib  # noqa: E402


@contextlib.contextmanager
def accuracy_on_device():
    """inside the block ``get_multi_acc`` returns device tensors instead of Python floats (no host round
    trip: required while a CUDA graph is being captured and wanted in the replayed / eager step); the
    reference behaviour (``SYNC_ACCURACY``) is restored on exit"""
    global SYNC_ACCURACY
    old = SYNC_ACCURACY
    SYNC_ACCURACY = False
    try:
        yield
    finally:
        SYNC_ACCURACY = old


def _parallel(*thunks):
    """independent passes through the same discriminator on parallel streams; module state they
    share (BatchNorm running statistics, spectral-norm u / v) is still updated in call order"""
    from cpcsv_b200 import streams
    return streams.concurrently(*thunks, enabled=PARALLEL_PASSES and not cfg.CASCADE_MODEL)


def _call(module, *inputs):
    return module(*inputs)


def get_multi_acc(predict, real):
    """fraction of positive labels whose sigmoid score is >= 0.5 (reference l.313-321)."""
    predict = torch.as_tensor(predict)
    real = torch.as_tensor(real)
    hit = ((torch.sigmoid(predict) >= 0.5) & (real == 1)).sum()
    acc = hit.float() / real.sum()
    return float(acc) if SYNC_ACCURACY else acc


def compute_discriminator_loss(netD, real_imgs, fake_imgs, real_labels, fake_labels, real_catelabels,
                               conditions, gpus, real_features=None):
    """real / wrong / fake conditional BCE terms (+ character classification on the real
    features).  Returns (errD, errD_real, errD_wrong, errD_fake, acc, consistency).
    ``real_features`` (extension): ``netD(real_imgs)`` already computed by the caller -- the real pass
    does not depend on the generator, so the trainer issues it before the fakes exist."""
    if conditions is None:
        raise NotImplementedError("unconditional discriminators are unused by CP-CSV")
    if netD.get_uncond_logits is not None:
        raise NotImplementedError("unconditional logits are unused by CP-CSV (get_uncond_logits is None)")
    batch_size = real_imgs.size(0)
    cond = conditions.detach()
    if real_features is None:
        real_features, fake_features = _parallel(lambda: _call(netD, real_imgs),
                                                 lambda: _call(netD, fake_imgs.detach()))
    else:
        fake_features = _call(netD, fake_imgs.detach())
    head = netD.get_cond_logits
    errD_real, errD_wrong, errD_fake = _parallel(
        lambda: F.binary_cross_entropy(_call(head, real_features, cond), real_labels),
        lambda: F.binary_cross_entropy(_call(head, real_features[:batch_size - 1], cond[1:]), fake_labels[1:]),
        lambda: F.binary_cross_entropy(_call(head, fake_features, cond), fake_labels))
    errD = errD_real + (errD_fake + errD_wrong) * 0.5
    acc = 0
    if netD.cate_classify is not None:
        cate_logits = _call(netD.cate_classify, real_features).squeeze()
        errD = errD + 1.0 * F.multilabel_soft_margin_loss(cate_logits, real_catelabels)
        acc = get_multi_acc(cate_logits.detach(), real_catelabels)
    consistency = 0
    if netD.seq_consisten_model is not None:
        # order-consistency critic (reference l.110-122): real stories, about half of them with their frames
        # permuted (label 1), BCE-with-logits on the critic's order logit
        shuffled, order_labels = create_random_shuffle(real_imgs)
        order_logits = _call(netD.seq_consisten_model, shuffled)
        consistency = F.binary_cross_entropy_with_logits(order_logits, order_labels.unsqueeze(-1))
        errD = errD + cfg.CONSISTENCY_RATIO * consistency
        consistency = consistency.detach()
    return errD, errD_real.detach(), errD_wrong.detach(), errD_fake.detach(), acc, consistency


def compute_generator_loss(netD, fake_imgs, real_imgs, real_labels, fake_catelabels, conditions, gpus):
    """BCE of D(fake) against the real label (+ classification).  Returns (err, acc, consistency)."""
    if conditions is None:
        raise NotImplementedError("unconditional discriminators are unused by CP-CSV")
    cond = conditions.detach()
    fake_features = _call(netD, fake_imgs)
    errD_fake = F.binary_cross_entropy(_call(netD.get_cond_logits, fake_features, cond), real_labels)
    acc = 0
    if netD.cate_classify is not None:
        cate_logits = _call(netD.cate_classify, fake_features).squeeze()
        errD_fake = errD_fake + 1.0 * F.multilabel_soft_margin_loss(cate_logits, fake_catelabels)
        acc = get_multi_acc(cate_logits.detach(), fake_catelabels)
    consistency = 0
    if netD.seq_consisten_model is not None:
        # reference l.155-169: the critic's logit of the generated story regressed onto its logit of the real one
        with torch.no_grad():
            real_logits = _call(netD.seq_consisten_model, real_imgs)
        fake_logits = _call(netD.seq_consisten_model, fake_imgs)
        consistency = F.mse_loss(fake_logits, real_logits)
        errD_fake = errD_fake + cfg.CONSISTENCY_RATIO * consistency
        consistency = consistency.detach()
    return errD_fake, acc, consistency


def KL_loss(mu, logvar):
    """-0.5 * mean(1 + logvar - mu^2 - exp(logvar))"""
    return -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())


def weights_init(m):
    """N(0, 0.02) for Conv / Linear weights, N(1, 0.02) for BatchNorm scale, zero biases;
    dispatch on the class name like the reference."""
    name = m.__class__.__name__
    if any(k in name for k in ("Conv", "BatchNorm", "Linear")):
        # `.data` writes bypass Tensor._version: tell the packed-operand cache (no-op before first use)
        from cpcsv_b200 import nets as _knets
        _knets.invalidate_weight_cache()
    if "Conv" in name:
        m.weight.data.normal_(0.0, 0.02)
    elif "BatchNorm" in name:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)
    elif "Linear" in name:
        m.weight.data.normal_(0.0, 0.02)
        if m.bias is not None:
            m.bias.data.fill_(0.0)


def count_param(model):
    return sum(p.numel() for p in model.parameters())


def mkdir_p(path):
    os.makedirs(path, exist_ok=True)


def save_model(netG, netD_im, netD_st, netD_se, epoch, model_dir, whole=False):
    """state_dict checkpoints with the reference's file names (l.323-338)."""
    if whole:
        torch.save(netG, "%s/netG.pkl" % model_dir)
        torch.save(netD_im, "%s/netD_im.pkl" % model_dir)
        torch.save(netD_st, "%s/netD_st.pkl" % model_dir)
        if netD_se is not None:
            torch.save(netD_se, "%s/netD_se.pkl" % model_dir)
        return
    torch.save(netG.state_dict(), "%s/netG_epoch_%d.pth" % (model_dir, epoch))
    torch.save(netD_im.state_dict(), "%s/netD_im_epoch_last.pth" % model_dir)
    torch.save(netD_st.state_dict(), "%s/netD_st_epoch_last.pth" % model_dir)
    if netD_se is not None:
        torch.save(netD_se.state_dict(), "%s/netD_se_epoch_last.pth" % model_dir)
