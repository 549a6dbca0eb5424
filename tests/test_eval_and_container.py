"""Eval driver (pad / crop / PSNR / bpp) and the bit-stream container (row 3)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F


def test_pad_crop_match_reference_arithmetic():
    """eval_model/__main__.py:103-115,126-128: centred zero padding to multiples of 64 and its inverse."""
    from compressai.utils.eval_model import crop, pad_to_multiple, psnr

    x = torch.rand(2, 3, 100, 130)
    xp, pads = pad_to_multiple(x)
    assert xp.shape == (2, 3, 128, 192) and pads == (31, 31, 14, 14)
    assert torch.equal(crop(xp, pads), x)
    assert float(xp[:, :, :14].abs().max()) == 0.0 and float(xp[:, :, :, :31].abs().max()) == 0.0
    x64 = torch.rand(1, 3, 64, 128)
    xp, pads = pad_to_multiple(x64)
    assert pads == (0, 0, 0, 0) and torch.equal(xp, x64)
    # odd remainders put the extra row / column at the bottom / right, like the reference
    _, pads = pad_to_multiple(torch.rand(1, 3, 61, 63))
    assert pads == (0, 1, 1, 2)
    a, b = torch.zeros(1, 3, 4, 4), torch.full((1, 3, 4, 4), 0.1)
    assert abs(psnr(a, b) - 20.0) < 1e-4
    assert abs(psnr(a, b) - (-10 * math.log10(F.mse_loss(a, b).item()))) < 1e-9


def test_container_round_trip_and_damage():
    from compressai.utils import container

    strings = [[b"\x01\x02\x03\x04" * 5, b"abcd" * 3], [b"zz" * 4, b""]]
    blob = container.pack("stf", strings, (12, 8), (700, 500), pads=(6, 6, 34, 34))
    got = container.unpack(blob)
    assert got == {"arch": "stf", "strings": strings, "shape": (12, 8), "image_size": (700, 500), "pads": (6, 6, 34, 34)}
    assert len(blob) == 4 + 2 + 3 + 8 + 8 + 8 + 4 + 16 + sum(len(s) for g in strings for s in g)
    for bad in (blob[:-1], blob + b"\0", b"XXXX" + blob[4:], blob[:20]):
        with pytest.raises(ValueError):
            container.unpack(bad)
    with pytest.raises(ValueError):
        container.pack("stf", [[b"a"], []], (1, 1), (64, 64))


def test_zoo_exports_the_reference_names():
    from compressai.zoo import load_state_dict, models

    assert set(models) >= {"stf", "cnn", "cnn2"}
    sd = load_state_dict({"module.g_a.0.weight": torch.zeros(1), "entropy_bottleneck._matrices.0": torch.zeros(1), "h_s.0.weight": torch.zeros(1)})
    assert set(sd) == {"g_a.0.weight", "entropy_bottleneck._matrix0"}


@pytest.mark.gpu
def test_eval_driver_end_to_end(tmp_path):
    """Images of a size that needs padding, through files on disk, a checkpoint on disk, real coding and entropy
    estimation, and the container: decode(unpack(pack(compress))) == the driver's reconstruction."""
    import numpy as np
    from PIL import Image

    from compressai.utils import container
    from compressai.utils.eval_model import collect_images, crop, eval_model, inference, inference_entropy_estimation, load_checkpoint, read_image
    from oracle import stf_ref, weights

    sd = weights.seeded_state_dict(stf_ref.template_state_dict(), seed=0, stress=True)
    ckpt = tmp_path / "ckpt.pth.tar"
    torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}}, ckpt)  # DataParallel-style keys, like a training run
    for i in range(2):
        img = (weights.seeded_image((1, 3, 100, 150), seed=60 + i)[0].permute(1, 2, 0).numpy() * 255 + 0.5).astype(np.uint8)
        Image.fromarray(img).save(tmp_path / f"im{i}.png")
    files = collect_images(str(tmp_path))
    assert [os.path.basename(f) for f in files] == ["im0.png", "im1.png"]
    net = load_checkpoint("stf", str(ckpt))
    x = read_image(files[0]).cuda()
    rv = inference(net, x)
    assert rv["x_hat"].shape == (1, 3, 100, 150) and list(rv["shape"]) == [2, 3]
    n_bytes = sum(len(s) for g in rv["strings"] for s in g)
    assert abs(rv["bpp"] - n_bytes * 8 / (100 * 150)) < 1e-9
    est = inference_entropy_estimation(net, x)
    assert abs(est["psnr"] - rv["psnr"]) < 1e-3          # decompress(compress(x)) == clamp(forward(x))
    # the stress weights put many symbols in the tails, which the estimate prices at the 1e-9 likelihood floor (30 bits)
    # while the coder's bypass path is cheaper: same order of magnitude only
    assert 0.5 < est["bpp"] / rv["bpp"] < 2.0
    blob = container.pack("stf", rv["strings"], rv["shape"], (100, 150), pads=(21, 21, 14, 14))
    got = container.unpack(blob)
    d = net.decompress(got["strings"], got["shape"])
    assert torch.equal(crop(d["x_hat"], got["pads"]), rv["x_hat"])
    res = eval_model(net, files, recon_path=str(tmp_path / "recon"))
    assert set(res) == {"psnr", "bpp", "encoding_time", "decoding_time"} and res["bpp"] > 0
    assert os.path.exists(tmp_path / "recon" / "im1.png")
