"""The ImageProcessor cases of image_processor.npz (shared by make_golden_image.py and the tests)."""
# (source stype, source shape, image_type, resize (w, h), normalize_type, trimming (top, left, bottom, right))
PROC_CASES = [
    ("RGB", (210, 160, 3), "GRAY_HW1", (84, 84), "0to1", None),         # InputImageBlockConfig "DQN" default (input_block.py:205)
    ("RGB", (210, 160, 3), "RGB", (96, 72), "0to1", None),              # "R2D3" default (:207)
    ("RGB", (96, 96, 3), "RGB", (96, 96), "0to1", None),                # "MuzeroAtari" default on a frame that already has the size
    ("RGB", (50, 70, 3), "GRAY_HW", (84, 84), "", None),                # upscale, uint8 out
    ("RGB", (64, 48, 3), "GRAY_HW1", (33, 77), "-1to1", (4, 6, 60, 40)),
    ("RGB", (40, 40, 3), "RGB", None, "-1to1", (-3, 5, 100, 31)),       # trimming clipped to the frame, no resize
    ("GRAY_HW", (64, 64), "RGB", (32, 32), "0to1", None),               # gray -> 3 equal channels
    ("GRAY_HW1", (30, 50, 1), "RGB", (25, 15), "", (2, 2, 28, 44)),
    ("GRAY_HW", (100, 120), "GRAY_HW1", (84, 84), "0to1", None),
    ("GRAY_HW1", (21, 34, 1), "GRAY_HW", (84, 84), "-1to1", None),
    ("RGB", (31, 17, 3), "GRAY_HW", None, "0to1", None),                # colour conversion + normalisation only
]

