"""The cases of tests/golden/prefilter.npz (shared by make_golden_prefilter.py and the tests; plain data)."""

FOCUS_CASES = (0, 1, 2, 3, 4, 7, 10, 13, 30, 31, 32, 33)
FILTERS = ("gaussian", "hamming")
# (case, entry, (out_w, out_h), geometry): roi = integer box | (angle, scale, tx, ty) of Affine2d.trs on top of the map of
# the whole frame onto the output
TENSOR_CASES = (
    (0, "crop", (64, 64), (20, 30, 420, 430)),
    (1, "crop", (48, 48), (-40, -25, 470, 485)),
    (2, "crop", (96, 64), (10, 5, 310, 235)),
    (4, "crop", (40, 40), (0, 0, 640, 480)),
    (5, "crop", (129, 129), (100, 100, 229, 229)),   # same size: no filter, plain copy
    (6, "crop", (129, 129), (120, 110, 200, 190)),   # up-scaling: the up-filter (linear) applies, not the down-filter
    (0, "affine", (64, 64), (0.35, 0.9, 3.0, -4.0)),
    (3, "affine", (96, 72), (-0.5, 0.8, 0.0, 6.0)),
    (4, "affine", (60, 60), (0.1, 1.0, -2.0, 1.0)),
)
