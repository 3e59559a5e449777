"""TEST INFRASTRUCTURE ONLY (never imported by the product path): CPU oracle of the training-side image front end.

The reference augments every training image on the CPU (torchlib/dataloader.py:138-217 ``create_albu_transform``):

    torchvision RandomAffine (PIL, NEAREST, fill 0)                      dataloader.py:139-145
    albumentations Resize(inference_resolution) = cv2.resize INTER_LINEAR on uint8   :147
    RandomCrop(train_resolution)                                         :148
    [p = albu_prob]  VerticalFlip(p), ..., GaussNoise(var_limit = noise_std**2, p = noise_prob)   :157-199
    ToFloat(255), Normalize(mean, std, max_pixel_value = 1)              :200-203

``reference_pipeline`` runs the REAL libraries that are present in this image (PIL's ImagingTransform through torchvision's
``F.affine``, OpenCV's ``cv2.resize``) with explicit random parameters; albumentations itself is not installed, so its
ToFloat / Normalize / VerticalFlip / RandomCrop -- index arithmetic and two float32 ops -- are restated from
albumentations 0.4.x ``functional.py``.  ``restated_pipeline`` is the integer restatement the CUDA kernel follows:

  * PIL ``affine_fixed`` (Geometry.c): 16.16 fixed point, xin = (FIX(c + a/2 + b/2) + FIX(a) x + FIX(b) y) >> 16, fill outside;
  * cv::resize 8U INTER_LINEAR (resize.cpp): 11-bit coefficients cvRound(f * 2048) from float32 fractions, the x fraction clamped
    at the borders, the y fraction NOT clamped (rows are clipped instead), horizontal pass in int32, vertical pass
    (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2; an exact 2x down-scale is the 2x2 box mean ((sum + 2) >> 2).

tests/test_oracle_augment.py pins the restatement to the libraries bit for bit; tests/test_augment_gpu.py holds the CUDA kernel
to the restatement."""
import math

import numpy as np

COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS


# ------------------------------------------------------------------------------------------------ parameters
def inverse_affine_matrix(center, angle, translate, scale, shear):
    """torchvision.transforms.functional._get_inverse_affine_matrix (the matrix RandomAffine hands to PIL), restated:
    M = T * C * RSS * C^-1 with RSS = R(angle) * Shear(sx, sy) * scale; returns the 6 coefficients of M^-1 (output -> input)."""
    rot = math.radians(angle)
    sx, sy = math.radians(shear[0]), math.radians(shear[1])
    cx, cy = center
    tx, ty = translate
    a = math.cos(rot - sy) / math.cos(sy)
    b = -math.cos(rot - sy) * math.tan(sx) / math.cos(sy) - math.sin(rot)
    c = math.sin(rot - sy) / math.cos(sy)
    d = -math.sin(rot - sy) * math.tan(sx) / math.cos(sy) + math.cos(rot)
    m = [d, -b, 0.0, -c, a, 0.0]
    m = [x / scale for x in m]
    m[2] += m[0] * (-cx - tx) + m[1] * (-cy - ty)
    m[5] += m[3] * (-cx - tx) + m[4] * (-cy - ty)
    m[2] += cx
    m[5] += cy
    return m


def fix16(matrix):
    """PIL Geometry.c affine_fixed: FIX(v) = floor(v * 65536 + 0.5); the half-pixel offset is folded into the constant terms"""
    a, b, c, d, e, f = matrix
    fx = lambda v: int(math.floor(v * 65536.0 + 0.5))
    return [fx(a), fx(b), fx(c + a * 0.5 + b * 0.5), fx(d), fx(e), fx(f + d * 0.5 + e * 0.5)]


def resize_tables(ssize: int, dsize: int, axis: str):
    """cv::resize INTER_LINEAR 8U index / coefficient tables for one axis: (s0, s1, c0, c1), each [dsize]"""
    scale = 1.0 / (dsize / ssize)  # cv::resize: inv_scale = dsize / ssize (double), scale = 1 / inv_scale
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if axis == "x":  # the x fraction is clamped at both borders
        lo, hi = s < 0, s >= ssize - 1
        f[lo], s[lo] = 0, 0
        f[hi], s[hi] = 0, ssize - 1
    c0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_ONE)).astype(np.int64)  # saturate_cast<short> = cvRound (half to even)
    c1 = np.rint(f * np.float32(COEF_ONE)).astype(np.int64)
    s0 = np.clip(s, 0, ssize - 1)
    s1 = np.clip(s + 1, 0, ssize - 1)
    return s0.astype(np.int32), s1.astype(np.int32), c0.astype(np.int32), c1.astype(np.int32)


def is_area2(Hs, Ws, R):
    """cv::resize switches INTER_LINEAR to the 2x2 box mean when both scales are exactly 2"""
    return Hs == 2 * R and Ws == 2 * R


# ------------------------------------------------------------------------------------------------ restatement
def affine_nearest_fixed(src, matrix, fill=0):
    """src [H,W] or [H,W,C] uint8 -> same shape"""
    H, W = src.shape[:2]
    a0, a1, a2, a3, a4, a5 = fix16(matrix)
    ys, xs = np.mgrid[0:H, 0:W].astype(np.int64)
    xin = (a2 + a1 * ys + a0 * xs) >> 16
    yin = (a5 + a4 * ys + a3 * xs) >> 16
    ok = (xin >= 0) & (xin < W) & (yin >= 0) & (yin < H)
    out = np.full_like(src, fill)
    out[ok] = src[yin[ok], xin[ok]]
    return out


def resize_linear_u8(src, R):
    """src [H,W] or [H,W,C] uint8 -> [R,R(,C)] uint8, cv2.resize(..., (R, R), interpolation=cv2.INTER_LINEAR)"""
    H, W = src.shape[:2]
    S = src.astype(np.int64)
    if is_area2(H, W, R):
        return ((S[0::2, 0::2] + S[0::2, 1::2] + S[1::2, 0::2] + S[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx0, sx1, ax0, ax1 = resize_tables(W, R, "x")
    sy0, sy1, by0, by1 = resize_tables(H, R, "y")
    sh = (1, R) + (1,) * (src.ndim - 2)
    rows = S[:, sx0] * ax0.reshape(sh) + S[:, sx1] * ax1.reshape(sh)          # [H, R(,C)] scaled by 2048
    S0, S1 = rows[sy0], rows[sy1]
    shy = (R, 1) + (1,) * (src.ndim - 2)
    out = (((by0.reshape(shy).astype(np.int64) * (S0 >> 4)) >> 16) + ((by1.reshape(shy).astype(np.int64) * (S1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def clahe_u8(src, clip_limit=1.0, tiles=8):
    """cv2.createCLAHE(clipLimit, (tiles, tiles)).apply(src) for uint8 [H,W] with H, W divisible by ``tiles`` (clahe.cpp
    CLAHE_CalcLut_Body + CLAHE_Interpolation_Body): per-tile histogram, clip at max(int(clip * area / 256), 1), excess spread
    evenly (+1 every 256 // residual bins for the remainder), LUT = cvRound(cumsum * 255 / area) with float32 products, then
    every pixel blends the four neighbouring tiles' LUTs bilinearly in float32 (separately rounded products) and cvRounds."""
    H, W = src.shape
    assert H % tiles == 0 and W % tiles == 0, "the reflect-padded case of CLAHE_Impl::apply is not restated"
    th, tw = H // tiles, W // tiles
    area = th * tw
    lut_scale = np.float32(255.0) / np.float32(area)
    clip = max(int(clip_limit * area / 256), 1) if clip_limit > 0 else 0
    luts = np.zeros((tiles, tiles, 256), dtype=np.uint8)
    for ty in range(tiles):
        for tx in range(tiles):
            h = np.bincount(src[ty * th:(ty + 1) * th, tx * tw:(tx + 1) * tw].reshape(-1), minlength=256).astype(np.int64)
            if clip > 0:
                clipped = int(np.maximum(h - clip, 0).sum())
                h = np.minimum(h, clip)
                batch = clipped // 256
                residual = clipped - batch * 256
                h += batch
                if residual:
                    step = max(256 // residual, 1)
                    i = 0
                    while i < 256 and residual > 0:
                        h[i] += 1
                        i += step
                        residual -= 1
            luts[ty, tx] = np.clip(np.rint(np.cumsum(h).astype(np.float32) * lut_scale), 0, 255).astype(np.uint8)
    f32 = np.float32

    def axis(n, t):
        pf = np.arange(n, dtype=np.float32) * (f32(1.0) / f32(t)) - f32(0.5)
        p1 = np.floor(pf).astype(np.int64)
        a = (pf - p1.astype(np.float32)).astype(np.float32)
        return np.maximum(p1, 0), np.minimum(p1 + 1, tiles - 1), a, (f32(1.0) - a).astype(np.float32)

    tx1, tx2, xa, xa1 = axis(W, tw)
    ty1, ty2, ya, ya1 = axis(H, th)
    v = src.astype(np.int64)
    L = luts.astype(np.float32)
    l11, l12 = L[ty1[:, None], tx1[None, :], v], L[ty1[:, None], tx2[None, :], v]
    l21, l22 = L[ty2[:, None], tx1[None, :], v], L[ty2[:, None], tx2[None, :], v]
    top = ((l11 * xa1[None, :]) + (l12 * xa[None, :])) * ya1[:, None]
    bot = ((l21 * xa1[None, :]) + (l22 * xa[None, :])) * ya[:, None]
    return np.clip(np.rint(top + bot), 0, 255).astype(np.uint8)


def clahe_reference(img, clip_limit=1.0, tiles=8):
    """albumentations.augmentations.functional.clahe (0.4.x) through OpenCV itself: grey images directly, 3-channel images on the
    L plane of an 8-bit RGB -> LAB -> RGB round trip"""
    import cv2

    mat = cv2.createCLAHE(clipLimit=clip_limit, tileGridSize=(tiles, tiles))
    if img.ndim == 2 or img.shape[2] == 1:
        return mat.apply(img)
    lab = cv2.cvtColor(img, cv2.COLOR_RGB2LAB)
    lab[:, :, 0] = mat.apply(lab[:, :, 0])
    return cv2.cvtColor(lab, cv2.COLOR_LAB2RGB)


def clahe_grey_rgb(v, grey_to_l, l_to_rgb, clip_limit=1.0, tiles=8):
    """the 3-channel recipe for a GREY image (R = G = B = v): L = grey_to_l[v], CLAHE, (r, g, b) = l_to_rgb[L'] -- the two
    tables are OpenCV's own conversion tabulated on grey pixels (scripts/make_lab_tables.py)"""
    lp = clahe_u8(np.asarray(grey_to_l, dtype=np.uint8)[v], clip_limit, tiles)
    return np.asarray(l_to_rgb, dtype=np.uint8).reshape(256, 3)[lp]


def to_float_normalize(u8, mean, std):
    """albumentations ToFloat(max_value=255) then Normalize(mean, std, max_pixel_value=1.0): float32 throughout,
    (x / 255 - mean) * reciprocal(std); HWC uint8 -> CHW float32 (AlbumentationsTorchTransform permutes, dataloader.py:50-51)"""
    x = u8.astype(np.float32) / np.float32(255.0)
    if x.ndim == 2:
        x = x[:, :, None]
    m = np.asarray(mean, dtype=np.float32)[: x.shape[2]]
    r = np.reciprocal(np.asarray(std, dtype=np.float32)[: x.shape[2]], dtype=np.float32)
    x = (x - m) * r
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def restated_pipeline(src, matrix, R, T, cy, cx, vflip, mean, std, clahe=False, rgb_tables=None):
    """deterministic part of create_albu_transform with explicit random parameters; src uint8 [H,W] or [H,W,C].
    clahe: dataloader.py:150-156 (after the crop, before the flip); rgb_tables = (grey_to_l, l_to_rgb) makes a grey source a
    3-channel image first (the RGB loader) and runs the LAB recipe"""
    w = affine_nearest_fixed(src, matrix)
    r = resize_linear_u8(w, R)
    c = r[cy:cy + T, cx:cx + T]
    if clahe:
        c = clahe_grey_rgb(c, *rgb_tables) if rgb_tables is not None else clahe_u8(np.ascontiguousarray(c))
    if vflip:
        c = c[::-1]
    return c, to_float_normalize(c, mean, std)


# ------------------------------------------------------------------------------------------------ the real libraries
def reference_pipeline(src, angle, translate, scale, shear, R, T, cy, cx, vflip, mean, std, clahe=False):
    """the same steps through torchvision/PIL and OpenCV themselves"""
    import cv2
    import torchvision.transforms.functional as TF
    from PIL import Image

    img = Image.fromarray(src)
    img = TF.affine(img, angle, list(translate), scale, list(shear))          # RandomAffine.forward, NEAREST, fill 0
    arr = np.array(img)
    arr = cv2.resize(arr, (R, R), interpolation=cv2.INTER_LINEAR)              # albumentations Resize
    arr = arr[cy:cy + T, cx:cx + T]                                            # RandomCrop with explicit offsets
    if clahe:
        arr = clahe_reference(np.ascontiguousarray(arr))                        # FromFloat(uint8) is the identity on uint8 input
    if vflip:
        arr = np.ascontiguousarray(arr[::-1, ...])                             # VerticalFlip = cv2.flip(img, 0)
    return arr, to_float_normalize(arr, mean, std)
