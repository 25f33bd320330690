"""Consumers of the generator's outputs, with the names, signatures and file layouts of the reference's
``miscc/utils.py`` (save_img_results l.205-228, images_to_numpy l.230-235, save_story_results
l.237-280, save_image_results l.282-301, save_all_img l.303-311, save_test_samples l.343-370,
save_train_samples l.372-400, inference_samples l.402-428, check_is_order / create_random_shuffle
l.17-44, compute_cyc_loss_* l.174-182), so that the reference's ``trainer.py`` / ``inference.py`` /
``main_*.py`` keep importing them from ``miscc.utils`` when this package replaces the reference's
files ("output layout consumers").  Host-side Python: tiling 64x64 frames into
uint8 sheets and writing PNG / npy / txt files -- nothing here is on the measured path.

Image sheets follow ``torchvision.utils.make_grid`` geometry (2-pixel black gutters, single-channel
images replicated to RGB) without depending on torchvision.
"""
import os
import random

import numpy as np
import torch

from miscc.config import cfg

GUTTER = 2      # torchvision.utils.make_grid default padding


def check_is_order(sequence):
    return bool((np.diff(sequence) >= 0).all())


def create_random_shuffle(stories, random_rate=0.5):
    """Order-consistency training data (only used with cfg.USE_SEQ_CONSISTENCY): with probability
    `random_rate` a story's frames are permuted into a non-sorted order and one frame may be replaced
    by the same-position frame of another story.  Returns (stories', labels) with label 1 = shuffled."""
    device = stories.device
    host = stories.cpu()
    n_stories, video_len = host.shape[0], host.shape[2]
    out, labels = [], []
    for idx in range(n_stories):
        shuffled = random_rate > np.random.random()
        labels.append(1 if shuffled else 0)
        if not shuffled:
            out.append(host[idx].clone())
            continue
        order = random.sample(range(video_len), video_len)
        while check_is_order(order):
            np.random.shuffle(order)
        story = host[idx][:, list(order)].clone()
        other = random.randint(0, n_stories - 1)
        if other != idx:
            pos = random.sample(range(video_len), 1)
            story[:, pos] = host[other][:, pos].clone()
        out.append(story)
    return torch.stack(out, 0).to(device), torch.tensor(labels, dtype=torch.float32, device=device)


def compute_cyc_loss_img(loss_fn, st_cyc_imgs, st_real_imgs):
    return loss_fn(st_cyc_imgs, st_real_imgs)


def compute_cyc_loss_txt(loss_fn, st_motion_cyc, st_motion_input):
    return loss_fn(st_motion_cyc, st_motion_input).mean()


# ------------------------------------------------------------------------------ image sheets
def _sheet(images, per_row):
    """[n, C, H, W] (or a list of [C, H, W]) -> [3 or C, rows*(H+g)+g, cols*(W+g)+g] with `per_row`
    cells per row and g = GUTTER black pixels around every cell; a single image comes back as is."""
    if isinstance(images, (list, tuple)):
        images = torch.stack(list(images), 0)
    if images.dim() == 3:
        images = images.unsqueeze(0)
    if images.shape[1] == 1:
        images = images.expand(-1, 3, -1, -1)
    n, c, h, w = images.shape
    if n == 1:
        return images[0]
    cols = min(per_row, n)
    rows = -(-n // cols)
    sheet = images.new_zeros((c, rows * (h + GUTTER) + GUTTER, cols * (w + GUTTER) + GUTTER))
    for k in range(n):
        y0 = (k // cols) * (h + GUTTER) + GUTTER
        x0 = (k % cols) * (w + GUTTER) + GUTTER
        sheet[:, y0:y0 + h, x0:x0 + w] = images[k]
    return sheet


def images_to_numpy(tensor):
    """[C, H, W] in [-1, 1] -> uint8 [H, W, C] (values outside the range are clipped).  A CUDA tensor is
    converted on the device (cpcsv_images_to_u8) and read back as uint8 HWC: a quarter of the bytes of the
    reference's fp32 read-back (miscc/utils.py:230-235), no host-side transpose / clip / cast."""
    tensor = tensor.detach()
    if tensor.is_cuda and tensor.dtype == torch.float32 and tensor.dim() == 3:
        from cpcsv_b200 import ops
        out = torch.empty((tensor.shape[1], tensor.shape[2], tensor.shape[0]), dtype=torch.uint8, device=tensor.device)
        ops.images_to_u8(tensor, out)
        return out.cpu().numpy()
    arr = tensor.cpu().numpy().transpose(1, 2, 0)
    return ((np.clip(arr, -1.0, 1.0) + 1.0) / 2.0 * 255.0).astype("uint8")


def _story_sheet(stories):
    """(B, C, V, H, W) -> one row of V frames per story, stories stacked vertically"""
    rows = [_sheet(stories[i].transpose(0, 1), cfg.VIDEO_LEN) for i in range(stories.shape[0])]
    return images_to_numpy(_sheet(rows, 1))


def save_story_results(ground_truth, images, texts, name, image_dir, step=0, lr=False):
    """uint8 sheet of the generated stories (ground truth appended to the right); the captions go to
    ``fake_samples_<name>.txt``.  Like the reference, the sheet itself is returned, not written."""
    sheet = _story_sheet(images)
    if ground_truth is not None:
        sheet = np.concatenate([sheet, _story_sheet(ground_truth)], axis=1)
    if texts is not None:
        with open("{}/fake_samples_{}.txt".format(image_dir, name), "w") as f:
            for idx in range(images.shape[0]):
                f.write(str(idx) + "--------------------------------------------------------\n")
                for frame_texts in texts:
                    f.write(frame_texts[idx] + "\n")
                f.write("\n\n")
    return sheet


def save_image_results(ground_truth, images, size=None):
    """per-frame results (N = ST_BATCH * V frames, e.g. segmentation masks) as a story sheet"""
    size = size if size is not None else cfg.IMSIZE
    shape = (cfg.TRAIN.ST_BATCH_SIZE, cfg.VIDEO_LEN, -1, size, size)

    def sheet_of(t):
        t = t.reshape(shape)
        return images_to_numpy(_sheet([_sheet(t[i], cfg.VIDEO_LEN) for i in range(t.shape[0])], 1))
    out = sheet_of(images)
    if ground_truth is not None:
        out = np.concatenate([out, sheet_of(ground_truth)], axis=1)
    return out


def _write_png(image, path, normalize=False):
    """torchvision.utils.save_image semantics: [0, 1] -> uint8 with rounding; values outside are
    clipped (the reference passes [-1, 1] frames un-normalised to save_all_img, so negative pixels
    come out black -- kept); normalize=True rescales by the tensor's min / max first."""
    import PIL.Image
    t = image.detach().float().cpu()
    if normalize:
        lo, hi = float(t.min()), float(t.max())
        t = (t - lo) / max(hi - lo, 1e-5)
    if t.dim() == 3 and t.shape[0] == 1:
        t = t.expand(3, -1, -1)
    arr = t.mul(255).add(0.5).clamp(0, 255).permute(1, 2, 0).to(torch.uint8).numpy()
    PIL.Image.fromarray(arr).save(path)


def save_all_img(images, count, image_dir):
    """every frame of (B, C, V, H, W) as <count+1>.png, <count+2>.png, ...; returns the new count"""
    for b in range(images.shape[0]):
        frames = images[b].transpose(0, 1)
        for i in range(frames.shape[0]):
            count += 1
            _write_png(frames[i], os.path.join(image_dir, "{}.png".format(count)))
    return count


def save_img_results(data_img, fake, texts, epoch, image_dir):
    num = cfg.VIS_COUNT
    fake = fake[0:num]
    if data_img is not None:
        _write_png(_sheet(data_img[0:num], 8), "%s/real_samples_epoch_%03d.png" % (image_dir, epoch), True)
        _write_png(_sheet(fake.detach(), 8), "%s/fake_samples_epoch_%03d.png" % (image_dir, epoch), True)
    else:
        _write_png(_sheet(fake.detach(), 8), "%s/lr_fake_samples_epoch_%03d.png" % (image_dir, epoch), True)
    if texts is not None:
        with open("%s/lr_fake_samples_epoch_%03d.txt" % (image_dir, epoch), "w") as f:
            for i in range(min(num, len(texts))):
                f.write(str(i) + ":" + texts[i] + "\n")


# ------------------------------------------------------------------------------ sampling loops
def _story_inputs(batch, device):
    """test / train story batch -> (real images, motion_input with labels appended, content_input)"""
    T = cfg.TEXT.DIMENSION
    real = batch["images"]
    desc = batch["description"][:, :, :T].to(device)
    labels = batch["labels"].to(device)
    return real, torch.cat((desc, labels), 2), desc, labels


def _generator_device(netG):
    return next(netG.parameters()).device


def _sample_loop(netG, dataloader, save_path, name_format):
    images, labels = [], []
    dev = _generator_device(netG)
    for i, batch in enumerate(dataloader):
        real, motion_input, content_input, catelabel = _story_inputs(batch, dev)
        _, fake, _, _, _, _, _ = netG.sample_videos(motion_input, content_input)
        save_story_results(real, fake, batch.get("text"), name_format.format(i), save_path)
        images.append(fake.detach().cpu().numpy())
        labels.append(catelabel.detach().cpu().numpy())
    np.save(save_path + "/images.npy", np.concatenate(images, 0))
    np.save(save_path + "/labels.npy", np.concatenate(labels, 0))


def save_test_samples(netG, dataloader, save_path):
    print("Generating Test Samples...")
    _sample_loop(netG, dataloader, save_path, "{:03d}")


def save_train_samples(netG, dataloader, save_path):
    print("Generating Train Samples...")
    _sample_loop(netG, dataloader, save_path, "{:05d}")


def inference_samples(netG, dataloader, save_path):
    """generated frames to <save_path>/<k>.png, the real ones to ./Evaluation/ref/<k>.png"""
    print("Generate and save images...")
    os.makedirs(save_path, exist_ok=True)
    os.makedirs("./Evaluation/ref", exist_ok=True)
    dev = _generator_device(netG)
    cnt_gen = cnt_ref = 0
    for batch in dataloader:
        real, motion_input, content_input, _ = _story_inputs(batch, dev)
        _, fake, _, _, _, _, _ = netG.sample_videos(motion_input, content_input)
        cnt_gen = save_all_img(fake, cnt_gen, save_path)
        cnt_ref = save_all_img(real, cnt_ref, "./Evaluation/ref")
