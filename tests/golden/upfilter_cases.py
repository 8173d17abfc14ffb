"""The cases of tests/golden/upfilter.npz (shared by make_golden_upfilter.py and the tests; plain data)."""

FILTERS = ("cubic", "lanczos")
FOCUS_CASES = (31, 32, 0, 1)  # tiny face (resize grows), tiny face rotated (the warp grows), and two that shrink (up-filter idle)
# (case, entry, (out_w, out_h), geometry): roi = integer box | (angle, scale, tx, ty) of Affine2d.trs on top of the map of
# the whole frame onto the output
TENSOR_CASES = (
    (3, "crop", (129, 129), (60, 50, 140, 130)),
    (3, "crop", (200, 160), (-10, -8, 90, 70)),      # box over the frame border, non-square
    (2, "crop", (129, 129), (100, 80, 229, 209)),    # same size: plain copy
    (3, "crop", (150, 150), (150, 120, 215, 185)),   # box over the right / bottom border
    (3, "affine", (300, 270), (0.3, 1.2, 5.0, -7.0)),
    (2, "affine", (400, 300), (-0.45, 1.05, 0.0, 3.0)),
    (0, "affine", (129, 129), (0.2, 4.0, -150.0, -180.0)),
)
