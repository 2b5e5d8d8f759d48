"""CPU restatement of the reference's image observation pipeline (TEST INFRASTRUCTURE; only tests/, smoke() and bench.py's CPU legs
may import this).

Follows srl/rl/processors/image_processor.py:104-154 (ImageProcessor.remap_observation: gray <-> colour, trimming, resize,
normalise, trailing channel axis).  The two pixel operations live in an un-vendored dependency, OpenCV (`cv2`, 4.13.0 in the build
container; the reference pins no version): their published fixed-point algorithms are restated here in numpy integer arithmetic --

  cv2.cvtColor(x, COLOR_RGB2GRAY) on uint8  (modules/imgproc/src/color_rgb.simd.hpp, RGB2Gray<uchar>): 15-bit coefficients
      y = (R * 9798 + G * 19235 + B * 3735 + 16384) >> 15
  cv2.resize(x, (w, h)) on uint8, INTER_LINEAR  (modules/imgproc/src/resize.cpp, resizeGeneric_ + HResizeLinear + VResizeLinear with
      FixedPtCast<int, uchar, 22>): 11-bit coefficients,
      horizontal: fx = (float)((dx + 0.5) * scale_x - 0.5), sx = floor(fx), fx -= sx; sx < 0 -> (0, fx = 0); sx >= W - 1 -> (W - 1, fx = 0)
                  row[dx] = S[sx] * round((1 - fx) * 2048) + S[sx + 1] * round(fx * 2048)
      vertical:   fy likewise but WITHOUT the border reset; the two source rows are clip(sy, 0, H - 1) and clip(sy + 1, 0, H - 1)
                  out = (((b0 * (row0 >> 4)) >> 16) + ((b1 * (row1 >> 4)) >> 16) + 2) >> 2

-- and pinned twice: against cv2 itself where it is importable (tests/test_image_oracle.py, skipped without cv2) and against
tests/golden/image_processor.npz, outputs of the reference's own ImageProcessor (tests/golden/make_golden_image.py).
"""
import numpy as np

GRAY_HW, GRAY_HW1, RGB = "GRAY_HW", "GRAY_HW1", "RGB"  # srl.base.define.SpaceTypes names


def rgb_to_gray(rgb: np.ndarray) -> np.ndarray:
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    return ((r * 9798 + g * 19235 + b * 3735 + 16384) >> 15).astype(np.uint8)


def linear_table(dst: int, src: int, border_reset: bool):
    """(index [dst] int32, coefficients [dst][2] int32) of one axis; float32 arithmetic where cv2 uses float."""
    scale = 1.0 / (float(dst) / float(src))
    idx = np.zeros(dst, np.int32)
    coef = np.zeros((dst, 2), np.int32)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if border_reset:
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= src - 1:
                s, f = src - 1, np.float32(0)
        idx[d] = s
        coef[d, 0] = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(2048)))
        coef[d, 1] = int(np.rint(f * np.float32(2048)))
    return idx, coef


def resize_linear_u8(img: np.ndarray, size_wh) -> np.ndarray:
    """cv2.resize(img, (w, h)) for uint8 [H, W] or [H, W, C]."""
    w, h = int(size_wh[0]), int(size_wh[1])
    H, W = img.shape[:2]
    squeeze = img.ndim == 2 or img.shape[2] == 1  # cv2 drops a single trailing channel
    x = img.astype(np.int64).reshape(H, W, -1)
    sx, ax = linear_table(w, W, True)
    sy, ay = linear_table(h, H, False)
    sx1 = np.minimum(sx + 1, W - 1)
    rows = x[:, sx, :] * ax[:, 0][None, :, None] + x[:, sx1, :] * ax[:, 1][None, :, None]
    r0, r1 = rows[np.clip(sy, 0, H - 1)], rows[np.clip(sy + 1, 0, H - 1)]
    b0, b1 = ay[:, 0][:, None, None].astype(np.int64), ay[:, 1][:, None, None].astype(np.int64)
    out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
    out = np.clip(out, 0, 255).astype(np.uint8)
    return out[..., 0] if squeeze else out


def clip_trimming(trimming, H, W):
    """image_processor.py:55-69"""
    top, left, bottom, right = trimming
    assert top < bottom and left < right
    return max(top, 0), max(left, 0), min(bottom, H), min(right, W)


def process(state: np.ndarray, prev_stype: str, image_type: str, resize=None, normalize_type: str = "", trimming=None, max_val=255):
    """ImageProcessor.remap_observation (image_processor.py:104-154) for a uint8 frame."""
    assert state.dtype == np.uint8
    if image_type == RGB and prev_stype in (GRAY_HW, GRAY_HW1):
        if prev_stype == GRAY_HW:
            state = state[..., np.newaxis]
        state = np.tile(state, (1, 1, 3))
    elif prev_stype == RGB and image_type in (GRAY_HW, GRAY_HW1):
        state = rgb_to_gray(state)
    if trimming is not None:
        top, left, bottom, right = clip_trimming(trimming, state.shape[0], state.shape[1])  # the space the reference clips against
        state = state[top:bottom, left:right]
    if resize is not None:
        state = resize_linear_u8(state, resize)
    if normalize_type == "0to1":
        state = state.astype(np.float32)
        state /= np.float32(max_val)
    elif normalize_type == "-1to1":
        state = state.astype(np.float32)
        state = (state * np.float32(2.0) / np.float32(max_val)) - np.float32(1.0)
    if state.ndim == 2 and image_type == GRAY_HW1:
        state = state[..., np.newaxis]
    return state
