"""Adaptive discriminator augmentation pipeline ("Training Generative Adversarial Networks with Limited Data") on this package's
ops: drop-in for the reference's `training/augment.py` `AugmentPipe` - same constructor arguments, buffers
(`p`, `Hz_geom`, `Hz_fbank`: a reference `augment_pipe` state dict loads unchanged) and `forward(images, debug_percentile=None)`.

What runs where: the geometric branch (augment.py:270-301) is reflect padding, a 2x `upfirdn2d.upsample2d` with the 12-tap sym6
low-pass, ONE `grid_sample_gradfix.grid_sample` of the composed inverse transform, and a 2x `upfirdn2d.downsample2d` - all on the
CUDA kernels of this package for CUDA tensors; the colour branch is a per-sample 4x4 matrix applied to the pixels; the optional
image-space band filter is a per-(sample, channel) separable FIR, issued as a grouped library convolution exactly like the
reference does (augment.py:398-399; `conv2d_gradfix` of this package covers groups = 1 only).

Every random decision draws the same tensors in the same order as the reference, so with equal seeds on the same device the two
produce the same images (tests/test_augment_cpu.py pins this against a fixture written by the reference itself).
"""
import numpy as np
import scipy.signal
import torch

from ..torch_utils.ops import grid_sample_gradfix
from ..torch_utils.ops import upfirdn2d

# low-pass decomposition filters of the two wavelets the pipeline uses (augment.py:20-37 lists more)
SYM2 = [-0.12940952255092145, 0.22414386804185735, 0.836516303737469, 0.48296291314469025]
SYM6 = [0.015404109327027373, 0.0034907120842174702, -0.11799011114819057, -0.048311742585633, 0.4910559419267466, 0.787641141030194,
        0.3379294217276218, -0.07263752278646252, -0.021060292512300564, 0.04472490177066578, 0.0017677118642428036,
        -0.007800708325034148]


def _mat(rows, device=None):
    """[..., R, C] matrix from nested rows of python scalars and / or equally shaped tensors (scalars are broadcast)"""
    tensors = [e for row in rows for e in row if torch.is_tensor(e)]
    if not tensors:
        return torch.tensor(rows, dtype=torch.float32, device=device)
    ref = tensors[0]
    cells = [e if torch.is_tensor(e) else torch.full(ref.shape, float(e), dtype=ref.dtype, device=ref.device) for row in rows for e in row]
    return torch.stack(cells, dim=-1).reshape(ref.shape + (len(rows), len(rows[0])))


def _shift2(tx, ty, **kw):
    return _mat([[1, 0, tx], [0, 1, ty], [0, 0, 1]], **kw)


def _zoom2(sx, sy, **kw):
    return _mat([[sx, 0, 0], [0, sy, 0], [0, 0, 1]], **kw)


def _turn2(angle):
    c, s = torch.cos(angle), torch.sin(angle)
    return _mat([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def _shift3(t):
    return _mat([[1, 0, 0, t], [0, 1, 0, t], [0, 0, 1, t], [0, 0, 0, 1]])


def _zoom3(s):
    return _mat([[s, 0, 0, 0], [0, s, 0, 0], [0, 0, s, 0], [0, 0, 0, 1]])


def _turn3(axis, angle):
    """rotation by `angle` about the unit vector axis[:3] (Rodrigues), homogeneous 4x4"""
    x, y, z = axis[0], axis[1], axis[2]
    s, c = torch.sin(angle), torch.cos(angle)
    k = 1 - c
    return _mat([[x * x * k + c, x * y * k - z * s, x * z * k + y * s, 0],
                 [y * x * k + z * s, y * y * k + c, y * z * k - x * s, 0],
                 [z * x * k - y * s, z * y * k + x * s, z * z * k + c, 0],
                 [0, 0, 0, 1]])


def _band_filter_bank():
    """4-band octave filter bank built from sym2 (augment.py:167-176): row i amplifies band i, the rows sum to an identity filter"""
    lo = np.asarray(SYM2)
    hi = lo * ((-1) ** np.arange(lo.size))
    lo2 = np.convolve(lo, lo[::-1]) / 2
    hi2 = np.convolve(hi, hi[::-1]) / 2
    bank = np.eye(4, 1)
    for i in range(1, bank.shape[0]):
        bank = np.dstack([bank, np.zeros_like(bank)]).reshape(bank.shape[0], -1)[:, :-1]       # zero-stuff: dilate by 2
        bank = scipy.signal.convolve(bank, [lo2])
        mid = bank.shape[1]
        bank[i, (mid - hi2.size) // 2: (mid + hi2.size) // 2] += hi2
    return torch.as_tensor(bank, dtype=torch.float32)


class AugmentPipe(torch.nn.Module):
    def __init__(self, xflip=0, rotate90=0, xint=0, xint_max=0.125,
                 scale=0, rotate=0, aniso=0, xfrac=0, scale_std=0.2, rotate_max=1, aniso_std=0.2, xfrac_std=0.125,
                 brightness=0, contrast=0, lumaflip=0, hue=0, saturation=0, brightness_std=0.2, contrast_std=0.5, hue_max=1, saturation_std=1,
                 imgfilter=0, imgfilter_bands=[1, 1, 1, 1], imgfilter_std=1,
                 noise=0, cutout=0, noise_std=0.1, cutout_size=0.5):
        super().__init__()
        given = dict(locals())
        for name, value in given.items():
            if name in ('self', 'given', '__class__'):
                continue
            setattr(self, name, list(value) if name == 'imgfilter_bands' else float(value))
        self.register_buffer('p', torch.ones([]))                               # overall probability multiplier (the ADA controller's knob)
        self.register_buffer('Hz_geom', upfirdn2d.setup_filter(SYM6))           # 12-tap separable low-pass of the resampling steps
        self.register_buffer('Hz_fbank', _band_filter_bank())

    # -- random parameters ----------------------------------------------------------------------------------------------------
    def _pick(self, value, gate_shape, prob, neutral, forced):
        """keep `value` where a fresh uniform draw of `gate_shape` falls below prob * p, else `neutral`; `forced` (debug mode,
        python scalar or 0-d tensor) overrides both.  The uniform draw happens AFTER `value` was drawn - the reference's order."""
        keep = torch.rand(gate_shape, device=value.device) < prob * self.p
        value = torch.where(keep, value, torch.full_like(value, neutral))
        return value if forced is None else torch.full_like(value, forced)

    def forward(self, images, debug_percentile=None):
        assert isinstance(images, torch.Tensor) and images.ndim == 4
        n, ch, height, width = images.shape
        dev = images.device
        dbg = None if debug_percentile is None else torch.as_tensor(debug_percentile, dtype=torch.float32, device=dev)
        uniform = lambda *shape: torch.rand(list(shape), device=dev)
        normal = lambda *shape: torch.randn(list(shape), device=dev)
        gauss_q = lambda: torch.erfinv(dbg * 2 - 1)                             # debug: the percentile's normal quantile / sqrt(2)

        # ---- inverse geometric transform G: output pixel -> input pixel (homogeneous, per sample) ----
        G = None

        def then(m):
            nonlocal G
            G = m if G is None else G @ m

        if self.xflip > 0:
            i = self._pick(torch.floor(uniform(n) * 2), [n], self.xflip, 0, None if dbg is None else torch.floor(dbg * 2))
            then(_zoom2(1 / (1 - 2 * i), 1 / torch.ones_like(i)))
        if self.rotate90 > 0:
            i = self._pick(torch.floor(uniform(n) * 4), [n], self.rotate90, 0, None if dbg is None else torch.floor(dbg * 4))
            then(_turn2(-(-np.pi / 2 * i)))
        if self.xint > 0:
            t = self._pick((uniform(n, 2) * 2 - 1) * self.xint_max, [n, 1], self.xint, 0, None if dbg is None else (dbg * 2 - 1) * self.xint_max)
            then(_shift2(-torch.round(t[:, 0] * width), -torch.round(t[:, 1] * height)))
        if self.scale > 0:
            s = self._pick(torch.exp2(normal(n) * self.scale_std), [n], self.scale, 1, None if dbg is None else torch.exp2(gauss_q() * self.scale_std))
            then(_zoom2(1 / s, 1 / s))
        # rotation is split into a pre- and a post-rotation around the anisotropic scaling, each with probability p_rot such that
        # P(pre or post) = rotate * p
        p_rot = 1 - torch.sqrt((1 - self.rotate * self.p).clamp(0, 1))
        if self.rotate > 0:
            a = (uniform(n) * 2 - 1) * np.pi * self.rotate_max
            a = torch.where(uniform(n) < p_rot, a, torch.zeros_like(a))
            if dbg is not None:
                a = torch.full_like(a, (dbg * 2 - 1) * np.pi * self.rotate_max)
            then(_turn2(-(-a)))
        if self.aniso > 0:
            s = self._pick(torch.exp2(normal(n) * self.aniso_std), [n], self.aniso, 1, None if dbg is None else torch.exp2(gauss_q() * self.aniso_std))
            then(_zoom2(1 / s, 1 / (1 / s)))
        if self.rotate > 0:
            a = (uniform(n) * 2 - 1) * np.pi * self.rotate_max
            a = torch.where(uniform(n) < p_rot, a, torch.zeros_like(a))
            if dbg is not None:
                a = torch.zeros_like(a)
            then(_turn2(-(-a)))
        if self.xfrac > 0:
            t = self._pick(normal(n, 2) * self.xfrac_std, [n, 1], self.xfrac, 0, None if dbg is None else gauss_q() * self.xfrac_std)
            then(_shift2(-(t[:, 0] * width), -(t[:, 1] * height)))

        if G is not None:
            images = self._resample(images, G)

        # ---- colour transform C: colour in -> colour out (homogeneous 4x4, per sample) ----
        C = None

        def before(m):
            nonlocal C
            C = m if C is None else m @ C

        luma = torch.tensor(np.asarray([1, 1, 1, 0]) / np.sqrt(3), dtype=torch.float32, device=dev)
        eye4 = torch.eye(4, device=dev)
        if self.brightness > 0:
            b = self._pick(normal(n) * self.brightness_std, [n], self.brightness, 0, None if dbg is None else gauss_q() * self.brightness_std)
            before(_shift3(b))
        if self.contrast > 0:
            c = self._pick(torch.exp2(normal(n) * self.contrast_std), [n], self.contrast, 1,
                           None if dbg is None else torch.exp2(gauss_q() * self.contrast_std))
            before(_zoom3(c))
        if self.lumaflip > 0:
            i = self._pick(torch.floor(uniform(n, 1, 1) * 2), [n, 1, 1], self.lumaflip, 0, None if dbg is None else torch.floor(dbg * 2))
            before(eye4 - 2 * torch.outer(luma, luma) * i)                      # Householder reflection about the luma axis
        if self.hue > 0 and ch > 1:
            a = self._pick((uniform(n) * 2 - 1) * np.pi * self.hue_max, [n], self.hue, 0, None if dbg is None else (dbg * 2 - 1) * np.pi * self.hue_max)
            before(_turn3(luma, a))
        if self.saturation > 0 and ch > 1:
            s = self._pick(torch.exp2(normal(n, 1, 1) * self.saturation_std), [n, 1, 1], self.saturation, 1,
                           None if dbg is None else torch.exp2(gauss_q() * self.saturation_std))
            ll = torch.outer(luma, luma)
            before(ll + (eye4 - ll) * s)
        if C is not None:
            if C.ndim == 2:
                C = C.unsqueeze(0)
            flat = images.reshape([n, ch, height * width])
            if ch == 3:
                flat = C[:, :3, :3] @ flat + C[:, :3, 3:]
            elif ch == 1:
                C = C[:, :3, :].mean(dim=1, keepdims=True)
                flat = flat * C[:, :, :3].sum(dim=2, keepdims=True) + C[:, :, 3:]
            else:
                raise ValueError('Image must be RGB (3 channels) or L (1 channel)')
            images = flat.reshape([n, ch, height, width])

        if self.imgfilter > 0:
            images = self._band_filter(images, dbg, gauss_q)

        # ---- corruptions ----
        if self.noise > 0:
            sigma = self._pick(normal(n, 1, 1, 1).abs() * self.noise_std, [n, 1, 1, 1], self.noise, 0,
                               None if dbg is None else torch.erfinv(dbg) * self.noise_std)
            images = images + normal(n, ch, height, width) * sigma
        if self.cutout > 0:
            size = torch.full([n, 2, 1, 1, 1], self.cutout_size, device=dev)
            size = torch.where(uniform(n, 1, 1, 1, 1) < self.cutout * self.p, size, torch.zeros_like(size))
            center = uniform(n, 2, 1, 1, 1)
            if dbg is not None:
                size = torch.full_like(size, self.cutout_size)
                center = torch.full_like(center, dbg)
            xs = torch.arange(width, device=dev).reshape([1, 1, 1, -1])
            ys = torch.arange(height, device=dev).reshape([1, 1, -1, 1])
            out_x = ((xs + 0.5) / width - center[:, 0]).abs() >= size[:, 0] / 2
            out_y = ((ys + 0.5) / height - center[:, 1]).abs() >= size[:, 1] / 2
            images = images * torch.logical_or(out_x, out_y).to(torch.float32)
        return images

    # -- geometric execution (augment.py:270-301) -------------------------------------------------------------------------------
    def _resample(self, images, G):
        n, ch, height, width = images.shape
        dev = images.device
        const = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
        # how far the transformed image corners reach outside the frame decides the reflect padding
        cx, cy = (width - 1) / 2, (height - 1) / 2
        corners = const([[-cx, -cy, 1], [cx, -cy, 1], [cx, cy, 1], [-cx, cy, 1]])
        moved = G @ corners.t()                                                 # [n, xyz, corner]
        taps = self.Hz_geom.shape[0] // 4
        reach = moved[:, :2, :].permute(1, 0, 2).flatten(1)                     # [xy, n * corner]
        reach = torch.cat([-reach, reach]).max(dim=1).values                    # x0, y0, x1, y1
        reach = reach + const([taps * 2 - cx, taps * 2 - cy] * 2)
        reach = reach.max(const([0, 0] * 2)).min(const([width - 1, height - 1] * 2))
        mx0, my0, mx1, my1 = reach.ceil().to(torch.int32)
        images = torch.nn.functional.pad(input=images, pad=[mx0, mx1, my0, my1], mode='reflect')
        G = _shift2((mx0 - mx1) / 2, (my0 - my1) / 2) @ G
        # 2x supersampling: the transform is conjugated into the upsampled pixel grid, then into affine_grid's [-1, 1] coordinates
        images = upfirdn2d.upsample2d(x=images, f=self.Hz_geom, up=2)
        G = _zoom2(2, 2, device=dev) @ G @ _zoom2(1 / 2, 1 / 2, device=dev)
        G = _shift2(-0.5, -0.5, device=dev) @ G @ _shift2(0.5, 0.5, device=dev)
        shape = [n, ch, (height + taps * 2) * 2, (width + taps * 2) * 2]
        G = _zoom2(2 / images.shape[3], 2 / images.shape[2], device=dev) @ G @ _zoom2(1 / (2 / shape[3]), 1 / (2 / shape[2]), device=dev)
        grid = torch.nn.functional.affine_grid(theta=G[:, :2, :], size=shape, align_corners=False)
        images = grid_sample_gradfix.grid_sample(images, grid)
        return upfirdn2d.downsample2d(x=images, f=self.Hz_geom, down=2, padding=-taps * 2, flip_filter=True)

    # -- image-space band filtering (augment.py:370-401) ----------------------------------------------------------------------
    def _band_filter(self, images, dbg, gauss_q):
        n, ch, height, width = images.shape
        dev = images.device
        bands = self.Hz_fbank.shape[0]
        assert len(self.imgfilter_bands) == bands
        power = torch.tensor(np.array([10, 1, 1, 1]) / 13, dtype=torch.float32, device=dev)     # expected 1/f power spectrum
        gain = torch.ones([n, bands], device=dev)
        for i, strength in enumerate(self.imgfilter_bands):
            t_i = torch.exp2(torch.randn([n], device=dev) * self.imgfilter_std)
            t_i = torch.where(torch.rand([n], device=dev) < self.imgfilter * self.p * strength, t_i, torch.ones_like(t_i))
            if dbg is not None:
                t_i = torch.full_like(t_i, torch.exp2(gauss_q() * self.imgfilter_std)) if strength > 0 else torch.ones_like(t_i)
            t = torch.ones([n, bands], device=dev)
            t[:, i] = t_i
            t = t / (power * t.square()).sum(dim=-1, keepdims=True).sqrt()      # keep the expected power unchanged
            gain = gain * t
        taps = (gain @ self.Hz_fbank).unsqueeze(1).repeat([1, ch, 1]).reshape([n * ch, 1, -1])
        pad = self.Hz_fbank.shape[1] // 2
        planes = images.reshape([1, n * ch, height, width])
        planes = torch.nn.functional.pad(input=planes, pad=[pad, pad, pad, pad], mode='reflect')
        # one filter per (sample, channel) plane: grouped library convolutions, as in the reference
        planes = torch.nn.functional.conv2d(planes, taps.unsqueeze(2), groups=n * ch)
        planes = torch.nn.functional.conv2d(planes, taps.unsqueeze(3), groups=n * ch)
        return planes.reshape([n, ch, height, width])
