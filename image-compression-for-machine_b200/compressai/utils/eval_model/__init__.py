"""Evaluation driver for the codecs: pad to a multiple of 64, compress / decompress (or
entropy estimation), crop, PSNR, bpp, timings; checkpoint loading.

Mirrors the image-compression part of /root/reference/compressai/utils/eval_model/__main__.py (`psnr` :75-77,
`inference` :96-139, `inference_entropy_estimation`, `load_checkpoint` :250-253, the per-file loop :442-487).  The
reference's detection / segmentation evaluators around it (COCO, detectron2, DeepLab) are out of scope.
All compute runs on the CUDA path of `compressai.models`; there is no CPU fallback.
"""
import math
import os
import time
from collections import defaultdict

import torch
import torch.nn.functional as F

IMG_EXTENSIONS = (".jpg", ".jpeg", ".png", ".ppm", ".bmp", ".pgm", ".tif", ".tiff", ".webp")


def collect_images(rootpath):
    """eval_model/__main__.py:67-72."""
    return sorted(os.path.join(rootpath, f) for f in os.listdir(rootpath) if os.path.splitext(f)[-1].lower() in IMG_EXTENSIONS)


def psnr(a, b):
    """eval_model/__main__.py:75-77."""
    mse = F.mse_loss(a.float(), b.float()).item()
    return -10 * math.log10(mse) if mse > 0 else float("inf")


def read_image(filepath):
    """RGB image file -> float tensor [3, H, W] in [0, 1] (eval_model/__main__.py:80-83)."""
    import numpy as np
    from PIL import Image

    assert os.path.isfile(filepath), filepath
    img = np.asarray(Image.open(filepath).convert("RGB"), dtype=np.uint8)
    return torch.from_numpy(img.copy()).permute(2, 0, 1).float().div_(255.0)


def pad_to_multiple(x, p=64):
    """Centred zero padding of [B, 3, H, W] to multiples of p (eval_model/__main__.py:103-115).  Returns (x_padded, pads)
    with pads = (left, right, top, bottom)."""
    h, w = x.size(2), x.size(3)
    new_h, new_w = (h + p - 1) // p * p, (w + p - 1) // p * p
    left = (new_w - w) // 2
    top = (new_h - h) // 2
    pads = (left, new_w - w - left, top, new_h - h - top)
    return F.pad(x, pads, mode="constant", value=0), pads


def crop(x, pads):
    """Undo pad_to_multiple (eval_model/__main__.py:126-128)."""
    left, right, top, bottom = pads
    return F.pad(x, (-left, -right, -top, -bottom))


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


@torch.no_grad()
def inference(model, x):
    """Real coding of a batch [B, 3, H, W] (or one image [3, H, W]): compress, decompress, crop, metrics per batch
    (eval_model/__main__.py:96-139).  bpp counts every byte of every string of the batch over its unpadded pixels."""
    if x.dim() == 3:
        x = x.unsqueeze(0)
    x_padded, pads = pad_to_multiple(x)
    _sync()
    start = time.time()
    out_enc = model.compress(x_padded)
    _sync()
    enc_time = time.time() - start
    start = time.time()
    out_dec = model.decompress(out_enc["strings"], out_enc["shape"])
    _sync()
    dec_time = time.time() - start
    x_hat = crop(out_dec["x_hat"], pads).clamp_(0, 1)
    num_pixels = x.size(0) * x.size(2) * x.size(3)
    bpp = sum(len(s) for group in out_enc["strings"] for s in group) * 8.0 / num_pixels
    return {"psnr": psnr(x, x_hat), "bpp": bpp, "encoding_time": enc_time, "decoding_time": dec_time, "x_hat": x_hat,
            "strings": out_enc["strings"], "shape": out_enc["shape"]}


@torch.no_grad()
def inference_entropy_estimation(model, x):
    """forward() only: bpp from the likelihoods, as the reference's estimation mode does."""
    if x.dim() == 3:
        x = x.unsqueeze(0)
    x_padded, pads = pad_to_multiple(x)
    _sync()
    start = time.time()
    out_net = model(x_padded)
    _sync()
    elapsed = time.time() - start
    x_hat = crop(out_net["x_hat"], pads).clamp_(0, 1)
    num_pixels = x.size(0) * x.size(2) * x.size(3)
    bpp = sum(float(torch.log(lk.float()).sum()) / (-math.log(2) * num_pixels) for lk in out_net["likelihoods"].values())
    return {"psnr": psnr(x, x_hat), "bpp": bpp, "encoding_time": elapsed / 2.0, "decoding_time": elapsed / 2.0, "x_hat": x_hat}


def load_checkpoint(arch, checkpoint_path, device="cuda"):
    """eval_model/__main__.py:250-253: checkpoint['state_dict'] -> zoo model -> eval (+ update(), :647-650)."""
    from compressai.zoo import load_state_dict, models

    ckpt = torch.load(checkpoint_path, map_location="cpu")
    state_dict = load_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt)
    net = models[arch].from_state_dict(state_dict).eval()
    net.update(force=True)
    return net.to(device)


def eval_model(model, filepaths, entropy_estimation=False, recon_path=None):
    """Per-file loop of the reference (:442-487): average psnr / bpp / times over the files."""
    device = next(model.parameters()).device
    metrics = defaultdict(float)
    for f in filepaths:
        x = read_image(f).to(device)
        rv = inference_entropy_estimation(model, x) if entropy_estimation else inference(model, x)
        if recon_path is not None:
            import numpy as np
            from PIL import Image

            os.makedirs(recon_path, exist_ok=True)
            img = (rv["x_hat"][0].permute(1, 2, 0).cpu().numpy() * 255.0 + 0.5).astype(np.uint8)
            Image.fromarray(img).save(os.path.join(recon_path, os.path.basename(f)))
        for k in ("psnr", "bpp", "encoding_time", "decoding_time"):
            metrics[k] += rv[k]
    return {k: v / max(len(filepaths), 1) for k, v in metrics.items()}
