"""StyleGAN2 discriminator on this package's layers: the caller of `conv2d_gradfix` (forward, data / weight gradients and the R1
double backward) in the training step of BASELINE config 4.  Same classes, constructor arguments and parameter names as the
reference (`training/networks.py:444-667`: DiscriminatorBlock, MinibatchStdLayer, DiscriminatorEpilogue, Discriminator), so the
`D` entry of a reference snapshot loads through `legacy.build_module`.

Every convolution is a `Conv2dLayer` of `training/synthesis.py`: with gradients enabled it goes through `conv2d_resample` ->
`conv2d_gradfix` (tcgen05 forward / dgrad / wgrad kernels, double-backward capable); under `torch.no_grad()` it takes the fused
single-launch route.  The minibatch-standard-deviation layer and the two fully connected layers are small library ops.
"""
import numpy as np
import torch

from ..torch_utils.ops import upfirdn2d
from .generator import MappingNetwork
from .synthesis import Conv2dLayer, FullyConnectedLayer

SQRT_HALF = float(np.sqrt(0.5))


class DiscriminatorBlock(torch.nn.Module):
    def __init__(self, in_channels, tmp_channels, out_channels, resolution, img_channels, first_layer_idx, architecture='resnet',
                 activation='lrelu', resample_filter=[1, 3, 3, 1], conv_clamp=None, use_fp16=False, fp16_channels_last=False,
                 freeze_layers=0):
        assert in_channels in [0, tmp_channels]
        assert architecture in ['orig', 'skip', 'resnet']
        super().__init__()
        self.in_channels, self.resolution, self.img_channels = in_channels, resolution, img_channels
        self.first_layer_idx, self.architecture, self.use_fp16 = first_layer_idx, architecture, use_fp16
        self.channels_last = use_fp16 and fp16_channels_last
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.num_layers = 0

        def add(name, *args, **kwargs):
            # Freeze-D (networks.py:476-481): the first `freeze_layers` layers hold their tensors as buffers
            trainable = self.first_layer_idx + self.num_layers >= freeze_layers
            self.num_layers += 1
            setattr(self, name, Conv2dLayer(*args, trainable=trainable, **kwargs))

        kw = dict(activation=activation, conv_clamp=conv_clamp)
        if in_channels == 0 or architecture == 'skip':
            add('fromrgb', img_channels, tmp_channels, kernel_size=1, **kw)
        add('conv0', tmp_channels, tmp_channels, kernel_size=3, **kw)
        add('conv1', tmp_channels, out_channels, kernel_size=3, down=2, resample_filter=resample_filter, **kw)
        if architecture == 'resnet':
            add('skip', tmp_channels, out_channels, kernel_size=1, bias=False, down=2, resample_filter=resample_filter)

    def forward(self, x, img, force_fp32=False, fused=True, impl='cuda'):
        dtype = torch.float16 if self.use_fp16 and not force_fp32 else torch.float32
        memory_format = torch.channels_last if self.channels_last and not force_fp32 else torch.contiguous_format
        kw = dict(fused=fused, impl=impl)
        if x is not None:
            assert tuple(x.shape[1:]) == (self.in_channels, self.resolution, self.resolution)
            x = x.to(dtype=dtype, memory_format=memory_format)
        if self.in_channels == 0 or self.architecture == 'skip':
            assert tuple(img.shape[1:]) == (self.img_channels, self.resolution, self.resolution)
            img = img.to(dtype=dtype, memory_format=memory_format)
            y = self.fromrgb(img, **kw)
            x = x + y if x is not None else y
            img = upfirdn2d.downsample2d(img, self.resample_filter, impl=impl) if self.architecture == 'skip' else None
        if self.architecture == 'resnet':
            y = self.skip(x, gain=SQRT_HALF, **kw)
            x = self.conv1(self.conv0(x, **kw), gain=SQRT_HALF, **kw)
            x = y.add_(x)
        else:
            x = self.conv1(self.conv0(x, **kw), **kw)
        assert x.dtype == dtype
        return x, img


class MinibatchStdLayer(torch.nn.Module):
    """appends, per group of `group_size` samples, the average standard deviation over the group as extra channel(s)"""

    def __init__(self, group_size, num_channels=1):
        super().__init__()
        self.group_size, self.num_channels = group_size, num_channels

    def forward(self, x):
        n, c, h, w = x.shape
        g = min(int(self.group_size), n) if self.group_size is not None else n
        f = self.num_channels
        y = x.reshape(g, -1, f, c // f, h, w)
        y = y - y.mean(dim=0)
        y = (y.square().mean(dim=0) + 1e-8).sqrt()
        y = y.mean(dim=[2, 3, 4]).reshape(-1, f, 1, 1).repeat(g, 1, h, w)
        return torch.cat([x, y], dim=1)


class DiscriminatorEpilogue(torch.nn.Module):
    def __init__(self, in_channels, cmap_dim, resolution, img_channels, architecture='resnet', mbstd_group_size=4, mbstd_num_channels=1,
                 activation='lrelu', conv_clamp=None):
        assert architecture in ['orig', 'skip', 'resnet']
        super().__init__()
        self.in_channels, self.cmap_dim, self.resolution, self.img_channels, self.architecture = \
            in_channels, cmap_dim, resolution, img_channels, architecture
        if architecture == 'skip':
            self.fromrgb = Conv2dLayer(img_channels, in_channels, kernel_size=1, activation=activation)
        self.mbstd = MinibatchStdLayer(group_size=mbstd_group_size, num_channels=mbstd_num_channels) if mbstd_num_channels > 0 else None
        self.conv = Conv2dLayer(in_channels + mbstd_num_channels, in_channels, kernel_size=3, activation=activation, conv_clamp=conv_clamp)
        self.fc = FullyConnectedLayer(in_channels * (resolution ** 2), in_channels, activation=activation)
        self.out = FullyConnectedLayer(in_channels, 1 if cmap_dim == 0 else cmap_dim)

    def forward(self, x, img, cmap, force_fp32=False, fused=True, impl='cuda'):
        assert tuple(x.shape[1:]) == (self.in_channels, self.resolution, self.resolution)
        x = x.to(dtype=torch.float32, memory_format=torch.contiguous_format)
        if self.architecture == 'skip':
            x = x + self.fromrgb(img.to(torch.float32), fused=fused, impl=impl)
        if self.mbstd is not None:
            x = self.mbstd(x)
        x = self.conv(x, fused=fused, impl=impl)
        x = self.out(self.fc(x.flatten(1), impl=impl), impl=impl)
        if self.cmap_dim > 0:
            assert tuple(cmap.shape[1:]) == (self.cmap_dim,)
            x = (x * cmap).sum(dim=1, keepdim=True) * (1 / np.sqrt(self.cmap_dim))
        return x


class Discriminator(torch.nn.Module):
    def __init__(self, c_dim, img_resolution, img_channels, architecture='resnet', channel_base=32768, channel_max=512, num_fp16_res=0,
                 conv_clamp=None, cmap_dim=None, block_kwargs={}, mapping_kwargs={}, epilogue_kwargs={}):
        super().__init__()
        self.c_dim, self.img_resolution, self.img_channels = c_dim, img_resolution, img_channels
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.block_resolutions = [2 ** i for i in range(self.img_resolution_log2, 2, -1)]
        ch = {res: min(channel_base // res, channel_max) for res in self.block_resolutions + [4]}
        fp16_resolution = max(2 ** (self.img_resolution_log2 + 1 - num_fp16_res), 8)
        if cmap_dim is None:
            cmap_dim = ch[4]
        if c_dim == 0:
            cmap_dim = 0
        common = dict(img_channels=img_channels, architecture=architecture, conv_clamp=conv_clamp)
        layer_idx = 0
        for res in self.block_resolutions:
            block = DiscriminatorBlock(ch[res] if res < img_resolution else 0, ch[res], ch[res // 2], resolution=res,
                                       first_layer_idx=layer_idx, use_fp16=(res >= fp16_resolution), **block_kwargs, **common)
            setattr(self, f'b{res}', block)
            layer_idx += block.num_layers
        if c_dim > 0:
            self.mapping = MappingNetwork(z_dim=0, c_dim=c_dim, w_dim=cmap_dim, num_ws=None, w_avg_beta=None, **mapping_kwargs)
        self.b4 = DiscriminatorEpilogue(ch[4], cmap_dim=cmap_dim, resolution=4, **epilogue_kwargs, **common)

    def forward(self, img, c, fused=True, impl='cuda', **block_kwargs):
        x = None
        for res in self.block_resolutions:
            x, img = getattr(self, f'b{res}')(x, img, fused=fused, impl=impl, **block_kwargs)
        cmap = self.mapping(None, c, impl=impl) if self.c_dim > 0 else None
        return self.b4(x, img, cmap, fused=fused, impl=impl)
