"""GPU-side training augmentation: ``create_albu_transform`` (torchlib/dataloader.py:138-217) as one kernel launch per batch.

The reference augments image by image on the CPU (PIL RandomAffine, then albumentations Resize / RandomCrop / VerticalFlip /
GaussNoise / ToFloat / Normalize) inside the DataLoader workers.  Here the host only draws the random parameters -- with the
distributions of ``RandomAffine.get_params`` and albumentations' ``get_params`` -- and packs the raw uint8 images; the pixels
are produced on the hospital's GPU by ``pm_augment_batch_u8_f32`` (csrc/augment.cu), which reproduces PIL's and OpenCV's
fixed-point arithmetic bit for bit (tests/test_augment_gpu.py, against oracle/augment_oracle.py which is pinned to the libraries).

``clahe = yes`` (the shipped pneumonia configs; dataloader.py:150-156) is cv::CLAHE restated bit for bit: two more launches between
the crop and the flip / noise / normalise tail.  albumentations runs it on the L plane of an 8-bit RGB -> LAB -> RGB round trip for
3-channel images; that is served for GREY sources (X-rays through the RGB loader) with OpenCV's conversion tabulated on grey pixels
(_lab_tables.py) -- a true colour source with clahe raises.  Not built (asking for one raises): the optional transforms the shipped
configs leave off -- RandomGamma, Blur, ElasticTransform, ... (dataloader.py:159-198)."""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from .._lib import AugSample, PrimiaError, call, ptr, stream

UNSUPPORTED = ("randomgamma", "randombrightness", "blur", "elastic", "optical_distortion", "grid_distortion", "grid_shuffle",
               "hsv", "invert", "cutout", "shadow", "fog", "sun_flare", "solarize", "equalize", "grid_dropout")


def inverse_affine_matrix(center, angle, translate, scale, shear):
    """the output->input matrix torchvision's RandomAffine hands to PIL (transforms/functional.py _get_inverse_affine_matrix):
    inverse of T * C * R(angle) * Shear(sx, sy) * scale * C^-1"""
    rot, sx, sy = math.radians(angle), math.radians(shear[0]), math.radians(shear[1])
    cx, cy = center
    tx, ty = translate
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    m = [d / scale, -b / scale, 0.0, -c / scale, a / scale, 0.0]
    m[2] += m[0] * (-cx - tx) + m[1] * (-cy - ty) + cx
    m[5] += m[3] * (-cx - tx) + m[4] * (-cy - ty) + cy
    return m


def fix16(m):
    """PIL Geometry.c affine_fixed: 16.16 fixed point with the pixel-centre offset folded into the constant terms"""
    f = lambda v: int(math.floor(v * 65536.0 + 0.5))
    return [f(m[0]), f(m[1]), f(m[2] + m[0] * 0.5 + m[1] * 0.5), f(m[3]), f(m[4]), f(m[5] + m[3] * 0.5 + m[4] * 0.5)]


def resize_tables(Hs: int, Ws: int, R: int) -> np.ndarray:
    """cv::resize 8U INTER_LINEAR tables [8, R] int32: sx0 sx1 ax0 ax1 / sy0 sy1 by0 by1 (resize.cpp: float32 fractions,
    cvRound(f * 2048); the column fraction is clamped at the borders, the row fraction is not -- rows are clipped instead)"""
    out = np.empty((8, R), dtype=np.int32)
    for k, (n, clamp) in enumerate(((Ws, True), (Hs, False))):
        scale = 1.0 / (R / n)
        f = ((np.arange(R, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        f = (f - s.astype(np.float32)).astype(np.float32)
        if clamp:
            lo, hi = s < 0, s >= n - 1
            f[lo], s[lo] = 0, 0
            f[hi], s[hi] = 0, n - 1
        out[4 * k + 0] = np.clip(s, 0, n - 1)
        out[4 * k + 1] = np.clip(s + 1, 0, n - 1)
        out[4 * k + 2] = np.rint((np.float32(1.0) - f) * np.float32(2048.0))
        out[4 * k + 3] = np.rint(f * np.float32(2048.0))
    return out


class GpuAugment:
    """``create_albu_transform(args, mean, std)`` for a whole batch on ``device``.

    args: the reference's Arguments fields rotation, translate, scale, shear, inference_resolution, train_resolution, albu_prob,
    individual_albu_probs, noise_std, noise_prob, pretrained (3 output channels; otherwise 1)."""

    def __init__(self, args, mean, std, device="cuda:0", seed=None):
        on = [k for k in UNSUPPORTED if getattr(args, k, False)]
        if on:
            raise PrimiaError(f"augmentations not built on the GPU front end: {on} (torchlib/dataloader.py:150-198)")
        self.args, self.device = args, torch.device(device)
        self.R, self.T = int(args.inference_resolution), int(args.train_resolution)
        self.cout = 3 if getattr(args, "pretrained", False) else 1
        mean = np.asarray(mean, dtype=np.float32).reshape(-1)
        std = np.asarray(std, dtype=np.float32).reshape(-1)
        if mean.size < self.cout:
            mean, std = np.repeat(mean[:1], self.cout), np.repeat(std[:1], self.cout)
        self.mean = np.ascontiguousarray(mean[: self.cout])
        self.rstd = np.reciprocal(np.ascontiguousarray(std[: self.cout]), dtype=np.float32)   # albumentations normalize()
        self.rng = np.random.default_rng(seed)
        self._tables = {}
        self.clahe = bool(getattr(args, "clahe", False))
        self._lab = None
        if self.clahe and self.T % 8:
            raise PrimiaError("clahe: train_resolution must be divisible by the 8 x 8 tile grid (the reflect-padded case of cv::CLAHE is not built)")

    # ---- random parameters (host)
    def sample_params(self, Hs: int, Ws: int) -> dict:
        a, g = self.args, self.rng
        rot, tr, sc, sh = float(a.rotation), float(a.translate), float(a.scale), float(a.shear)
        # RandomAffine.get_params: angle ~ U(-rot, rot); t ~ round(U(-tr * size, tr * size)); scale ~ U(1 - s, 1 + s); x-shear ~ U(-sh, sh)
        p = {"angle": float(g.uniform(-rot, rot)),
             "translate": (int(round(g.uniform(-tr * Ws, tr * Ws))), int(round(g.uniform(-tr * Hs, tr * Hs)))) if tr > 0 else (0, 0),
             "scale": float(g.uniform(1.0 - sc, 1.0 + sc)), "shear": (float(g.uniform(-sh, sh)), 0.0)}
        # albumentations RandomCrop: y1 = int((H - h) * random()), x1 = int((W - w) * random())
        p["cy"], p["cx"] = int((self.R - self.T) * g.random()), int((self.R - self.T) * g.random())
        group = g.random() < float(getattr(a, "albu_prob", 1.0))                      # a.Compose(train_tf_albu, p=albu_prob)
        p["flip"] = bool(group and g.random() < float(getattr(a, "individual_albu_probs", 0.0)))
        p["noise_sigma"], p["noise_seed"] = 0.0, 0
        if group and g.random() < float(getattr(a, "noise_prob", 0.0)):                # GaussNoise(var_limit=noise_std**2)
            p["noise_sigma"] = math.sqrt(g.uniform(0.0, float(a.noise_std) ** 2))
            p["noise_seed"] = int(g.integers(1, 2 ** 63))
        return p

    # ---- pixels (device)
    def apply(self, images, params, return_u8=False):
        """images: list of uint8 arrays [H,W] or [H,W,C] (C in 1, 3); params: one dict per image (``sample_params``)"""
        B, R, T = len(images), self.R, self.T
        assert B == len(params) and B > 0
        descs = (AugSample * B)()
        offs, tabs, tab_index, total = [], [], {}, 0
        for i, (im, p) in enumerate(zip(images, params)):
            im = np.ascontiguousarray(im)
            if im.dtype != np.uint8 or im.ndim not in (2, 3) or (im.ndim == 3 and im.shape[2] not in (1, 3)):
                raise PrimiaError(f"image {i}: expected uint8 [H,W] or [H,W,1|3], got {im.dtype} {im.shape}")
            Hs, Ws = im.shape[:2]
            C = 1 if im.ndim == 2 else im.shape[2]
            if C > self.cout:
                raise PrimiaError("a 3-channel source needs pretrained = yes (3 output channels)")
            m = inverse_affine_matrix((Ws * 0.5, Hs * 0.5), p["angle"], p["translate"], p["scale"], p["shear"])
            d = descs[i]
            d.src_off, d.Hs, d.Ws, d.C = total, Hs, Ws, C
            for k, v in enumerate(fix16(m)):
                if not -2 ** 31 <= v < 2 ** 31:
                    raise PrimiaError("affine coefficients exceed PIL's 16.16 fixed-point range")
                d.fix[k] = v
            d.cy, d.cx, d.flip = int(p["cy"]), int(p["cx"]), int(bool(p["flip"]))
            d.area2 = int(Hs == 2 * R and Ws == 2 * R)
            d.noise_sigma, d.noise_seed = float(p.get("noise_sigma", 0.0)), int(p.get("noise_seed", 0))
            key = (Hs, Ws)
            if key not in tab_index:
                if key not in self._tables:
                    self._tables[key] = resize_tables(Hs, Ws, R)
                tab_index[key] = 8 * R * len(tabs)
                tabs.append(self._tables[key])
            d.tab_off = tab_index[key]
            offs.append(im)
            total += im.size
        host = torch.empty(total, dtype=torch.uint8).pin_memory() if self.device.type == "cuda" else torch.empty(total, dtype=torch.uint8)
        o = 0
        for im in offs:
            host[o:o + im.size] = torch.from_numpy(im.reshape(-1))
            o += im.size
        with torch.cuda.device(self.device):
            src = host.to(self.device, non_blocking=True)
            tables = torch.from_numpy(np.concatenate(tabs, axis=0).reshape(-1)).to(self.device)
            dbuf = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(self.device)
            out = torch.empty((B, self.cout, T, T), dtype=torch.float32, device=self.device)
            u8 = torch.empty((B, self.cout, T, T), dtype=torch.uint8, device=self.device) if return_u8 else None
            mean_p, rstd_p = self.mean.ctypes.data_as(ctypes.c_void_p), self.rstd.ctypes.data_as(ctypes.c_void_p)
            if not self.clahe:
                call("pm_augment_batch_u8_f32", ptr(src), ptr(dbuf), ptr(tables), B, R, T, self.cout, mean_p, rstd_p, ptr(out),
                     ptr(u8) if u8 is not None else None, stream())
            else:
                # crop (no flip, no noise) -> per-tile CLAHE tables -> blend + flip + noise + ToFloat + Normalize
                if self.cout == 3 and any(d.C != 1 for d in descs):
                    raise PrimiaError("clahe on a colour source needs the full 8-bit RGB <-> LAB conversion, which is not built; "
                                      "grey sources (X-rays through the RGB loader) are")
                plain = (AugSample * B)()
                ctypes.memmove(plain, descs, ctypes.sizeof(descs))
                for d in plain:
                    d.flip, d.noise_sigma = 0, 0.0
                pbuf = torch.frombuffer(bytearray(bytes(plain)), dtype=torch.uint8).to(self.device)
                crop = torch.empty((B, T, T), dtype=torch.uint8, device=self.device)
                call("pm_augment_batch_u8_f32", ptr(src), ptr(pbuf), ptr(tables), B, R, T, 1, mean_p, rstd_p, None, ptr(crop), stream())
                pre = post = None
                if self.cout == 3:
                    if self._lab is None:
                        from ._lab_tables import GREY_TO_L, L_TO_RGB

                        self._lab = (torch.frombuffer(bytearray(GREY_TO_L), dtype=torch.uint8).to(self.device),
                                     torch.frombuffer(bytearray(L_TO_RGB), dtype=torch.uint8).to(self.device))
                    pre, post = self._lab
                luts = torch.empty((B, 64, 256), dtype=torch.uint8, device=self.device)
                o = lambda t: ptr(t) if t is not None else None
                call("pm_clahe_luts_u8", ptr(crop), o(pre), B, T, 8, ctypes.c_float(1.0), ptr(luts), stream())   # clip_limit = (1, 1)
                call("pm_augment_clahe_finish_f32", ptr(crop), o(pre), ptr(luts), o(post), ptr(dbuf), B, T, 8, self.cout, mean_p, rstd_p,
                     ptr(out), ptr(u8) if u8 is not None else None, stream())
        return (out, u8) if return_u8 else out

    def __call__(self, images):
        return self.apply(images, [self.sample_params(*np.asarray(im).shape[:2]) for im in images])
