"""WACNN ("cnn") and WACNN2 ("cnn2", codec part) on the B200-native kernels.

Drop-in for /root/reference/compressai/models/cnn.py:23-332 (and the byte-identical codec of cnn2.py:26-133,255-377):
same constructor, parameter / buffer names and forward / compress / decompress / update / load_state_dict /
from_state_dict contract.  g_a / g_s are 5x5 stride-2 (transposed) convolutions, GDN / IGDN and two gated
window-attention blocks (cnn.py:31-52); every convolution, the GDN GEMM and the attention linears run on the
tcgen05 implicit-GEMM kernel (csrc/conv.cu), the rest in csrc/wacnn.cu; the hyper-prior, the 10-slice
channel-conditional loop and the rANS coders are shared with STF (models/_context.py).

H and W must be multiples of 64, as for the reference (its window attention does not pad, and its eval pads
inputs to multiples of 64).  The RetinaNet teacher / student of cnn2.py:135-253 are a downstream vision task and
are not part of this package; WACNN2.forward is broken in the reference as shipped.
"""
import math

import torch
import torch.nn as nn

from compressai._native import ACT_NONE, OUT_BF16, check, lib, stream_ptr
from compressai.entropy_models import EntropyBottleneck, GaussianConditional
from compressai.layers import GDN, Win_noShift_Attention, conv3x3, subpel_conv3x3

from ._context import ChannelContextCodec
from ._engine import Engine
from .utils import conv, deconv

SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def _context_stack(cin):
    chans, mods = (224, 176, 128, 64, 32), []
    for i, c in enumerate(chans):
        mods.append(conv(cin, c, stride=1, kernel_size=3))
        if i < len(chans) - 1:
            mods.append(nn.GELU())
        cin = c
    return nn.Sequential(*mods)


class WACNN(ChannelContextCodec):
    latent_channels = 320
    hyper_channels = 192
    num_slices = 10
    max_support_slices = 5
    scale_table_fn = staticmethod(get_scale_table)

    def __init__(self, N=192, M=320, **kwargs):
        super().__init__(**kwargs)
        if (N, M) != (192, 320):
            raise ValueError("the CUDA WACNN path is built for the reference configuration N=192, M=320")
        self.g_a = nn.Sequential(
            conv(3, N, kernel_size=5, stride=2), GDN(N), conv(N, N, kernel_size=5, stride=2), GDN(N),
            Win_noShift_Attention(dim=N, num_heads=8, window_size=8, shift_size=4),
            conv(N, N, kernel_size=5, stride=2), GDN(N), conv(N, M, kernel_size=5, stride=2),
            Win_noShift_Attention(dim=M, num_heads=8, window_size=4, shift_size=2))
        self.g_s = nn.Sequential(
            Win_noShift_Attention(dim=M, num_heads=8, window_size=4, shift_size=2),
            deconv(M, N, kernel_size=5, stride=2), GDN(N, inverse=True), deconv(N, N, kernel_size=5, stride=2), GDN(N, inverse=True),
            Win_noShift_Attention(dim=N, num_heads=8, window_size=8, shift_size=4),
            deconv(N, N, kernel_size=5, stride=2), GDN(N, inverse=True), deconv(N, 3, kernel_size=5, stride=2))
        G = nn.GELU
        self.h_a = nn.Sequential(conv3x3(320, 320), G(), conv3x3(320, 288), G(), conv3x3(288, 256, stride=2), G(),
                                 conv3x3(256, 224), G(), conv3x3(224, 192, stride=2))
        hs = lambda: nn.Sequential(conv3x3(192, 192), G(), subpel_conv3x3(192, 224, 2), G(), conv3x3(224, 256), G(),
                                   subpel_conv3x3(256, 288, 2), G(), conv3x3(288, 320))
        self.h_mean_s = hs()
        self.h_scale_s = hs()
        self.cc_mean_transforms = nn.ModuleList(_context_stack(320 + 32 * min(i, 5)) for i in range(10))
        self.cc_scale_transforms = nn.ModuleList(_context_stack(320 + 32 * min(i, 5)) for i in range(10))
        self.lrp_transforms = nn.ModuleList(_context_stack(320 + 32 * min(i + 1, 6)) for i in range(10))
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.gaussian_conditional = GaussianConditional(None)
        self._engine = Engine(self)

    @classmethod
    def from_state_dict(cls, state_dict):
        net = cls(192, 320)
        net.load_state_dict(state_dict)
        return net

    # ---------------------------------------------------------------------------------- transforms
    def _run_sequence(self, seq, t, B, h, w, last_f32=False):
        """Walk g_a / g_s over channels-last bf16 activations t [B*h*w, C]."""
        e = self._engine
        mods = list(seq)
        for k, m in enumerate(mods):
            if isinstance(m, nn.Conv2d):
                t = e.conv(t, B, h, w, e.packed(m))
                h, w = (h + 2 * m.padding[0] - m.kernel_size[0]) // m.stride[0] + 1, (w + 2 * m.padding[1] - m.kernel_size[1]) // m.stride[1] + 1
            elif isinstance(m, nn.ConvTranspose2d):
                t = e.conv(t, B, h, w, e.packed(m))
                h, w = 2 * h, 2 * w
            elif isinstance(m, GDN):
                t = e.gdn(t, B, h, w, m)
            elif isinstance(m, Win_noShift_Attention):
                t = e.gated_window_block(t, B, h, w, m, out_f32=last_f32 and k == len(mods) - 1)
            else:
                raise TypeError(type(m))
        return t, h, w

    def _analysis(self, x):
        """g_a (cnn.py:31-41): image [B,3,H,W] -> y fp32 channels-last [B*h*w, 320]."""
        B, _, H, W = x.shape
        x = x.detach().float().contiguous()
        t = torch.empty((B * H * W, 8), dtype=torch.bfloat16, device=x.device)
        check(lib().icm_image_to_nhwc(x.data_ptr(), t.data_ptr(), B, 3, H, W, 8, stream_ptr()), "icm_image_to_nhwc")
        return self._run_sequence(self.g_a, t, B, H, W, last_f32=True)

    def _synthesis(self, y_hat, B, h, w, clamp):
        """g_s (cnn.py:42-52, 331): y_hat fp32 channels-last [B*h*w, 320] -> x_hat [B,3,16h,16w]."""
        t, H, W = self._run_sequence(self.g_s, self._engine.cast_bf16(y_hat), B, h, w)
        out = torch.empty((B, 3, H, W), dtype=torch.float32, device=t.device)
        check(lib().icm_nhwc_to_image(t.data_ptr(), out.data_ptr(), B, 3, H, W, t.shape[1], 1 if clamp else 0, stream_ptr()), "icm_nhwc_to_image")
        return out


class WACNN2(WACNN):
    """The codec of cnn2.py (identical layers to WACNN, cnn2.py:34-133); the RetinaNet teacher/student heads of the
    reference class are out of scope."""
