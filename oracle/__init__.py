"""CPU oracle for the augmentation / label-transform hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package restates, in numpy (+ the same OpenCV calls the reference makes), the algorithm of
`trackertraincode/datatransformation` of opentrack/neuralnet-tracker-traincode.  Every function cites the
reference file:line it follows (paths relative to the reference checkout).

Rules:
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
    import this package -- and only as the checker or the timed CPU baseline, never as a product code path;
  * nothing in here may read `/root/reference` at run time (it does not exist on the GPU box).

Parity pinning:
  * geometric + label half: pinned against outputs of the unmodified reference run in the authoring
    container (`tests/golden/make_golden.py` -> `tests/golden/*.npz`) and against the reference's own
    known-answer vectors (`test/test_affine_img_trafo.py:49-69`);
  * OpenCV arithmetic (`oracle/cv2_model.py`): pinned bit-exact against the live cv2 binary (same wheel on
    the GPU box) in `tests/test_oracle_cv2_model.py`; the anti-alias down-filters and the cubic / Lanczos
    up-filters in `tests/test_oracle_prefilter.py` / `tests/test_oracle_upfilters.py` (against cv2 and against
    reference outputs, `tests/golden/prefilter.npz`, `upfilter.npz`);
  * head-model roi (`oracle/headmodel.py`): reference outputs with its real face model (`tests/golden/headroi.npz`);
  * evaluation-side back-transform chain, rotation labels: `tests/golden/backtransform.npz`, `perspective.npz`;
  * photometric half (kornia): **parity unpinned against kornia itself** -- kornia is not installed anywhere we
    can run and the reference has no test for it; `oracle/photometric.py` freezes the written specification of
    SURVEY.md 8(c) and is pinned against the torch primitives kornia calls (`oracle/photometric_torch.py`,
    `tests/test_oracle_photometric_torch.py`, `tests/test_gpu_photometric_torch.py`).
"""
