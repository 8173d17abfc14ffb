"""The reference's image/label consistency KATs (test/test_affine_img_trafo.py:95-216) against the CPU oracle: the oracle's
crops put the heat-map peaks on the oracle's transformed landmarks within the reference's tolerances."""
from functools import partial

import numpy as np
import pytest
import torch

from oracle import geometric as ogeo
from oracle.geometric import Sample
from test_gpu_reference_kats import centroids, make_test_data


@pytest.mark.parametrize("way, scale, angle, tr, extra_tol", [
    ("up", 1.0, 0.0, (0.0, 0.0), 0.0), ("down", 1.0, 0.0, (0.0, 0.0), 0.0),                    # test_scalingtrafo
    ("up", 0.75, 20.0 * np.pi / 180.0, (-0.1, 0.03), 0.5), ("down", 0.75, 20.0 * np.pi / 180.0, (-0.1, 0.03), 0.2),  # ..._with_randomizer
])
def test_heatmap_peaks_follow_landmarks(way, scale, angle, tr, extra_tol):
    S, R, img, points, roi = make_test_data(way)
    a32 = np.float32(angle)
    cs = (float(torch.cos(torch.tensor(a32))), float(torch.sin(torch.tensor(a32))))
    crops = []
    for c in range(3):
        s = Sample((S, S), {"image": img[c].numpy()[..., None], "pt3d_68": points.numpy(), "roi": roi.numpy()},
                   {"image": "img", "pt3d_68": "pts", "roi": "roi"})
        out, _ = ogeo.focus_roi(s, ogeo.RoiFocusParams(np.float32(scale), a32, np.asarray(tr, np.float32), cs), R)
        crops.append(torch.from_numpy(out.data["image"].reshape(R, R).astype(np.float64)))
    np.testing.assert_allclose(centroids(torch.stack(crops)).numpy(), out.data["pt3d_68"][:, :2], atol=0.01 + extra_tol)
