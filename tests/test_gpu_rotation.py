"""GPU parity of the rotation-label entry points (through the C ABI): PerspectiveCorrector.corrected_rotation /
_make_look_at_matrix (eval.py:485-544) and torchquaternion.tomatrix / from_matrix, against the oracle, the reference's golden
outputs and the reference's known-answer tests.  Tolerance: 1e-4 relative (north_star), observed ~1e-7."""
import math
import os

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

from oracle import perspective as P
from test_oracle_perspective import FOVS, KATS

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "perspective.npz"))


def _cuda(a):
    return torch.from_numpy(np.asarray(a)).cuda()


@pytest.mark.parametrize("fov", FOVS)
def test_corrected_rotation(fov):
    from trackertraincode_b200.eval import PerspectiveCorrector

    pc = PerspectiveCorrector(fov)
    shared = pc.corrected_rotation(torch.from_numpy(G["image_sizes"][0]), _cuda(G["coord"]), _cuda(G["pose"])).cpu().numpy()
    np.testing.assert_allclose(shared, G[f"corrected_shared_fov{fov:g}"], rtol=1e-4, atol=2e-6)
    np.testing.assert_allclose(shared, P.corrected_rotation(fov, G["image_sizes"][0], G["coord"], G["pose"]), rtol=1e-4, atol=2e-6)
    # [B,2] sizes: the reference divides by row 0 (eval.py:525) -- same quirk here and in the oracle
    per_b = pc.corrected_rotation(_cuda(G["image_sizes"]), _cuda(G["coord"]), _cuda(G["pose"])).cpu().numpy()
    np.testing.assert_allclose(per_b, P.corrected_rotation(fov, G["image_sizes"], G["coord"], G["pose"]), rtol=1e-4, atol=2e-6)
    # one sample at a time = the semantics of the reference's own test
    one = torch.stack([pc.corrected_rotation(torch.from_numpy(G["image_sizes"][i]), _cuda(G["coord"][i]), _cuda(G["pose"][i])) for i in range(16)])
    np.testing.assert_allclose(one.cpu().numpy(), G[f"corrected_fov{fov:g}"][:16], rtol=1e-4, atol=2e-6)


def test_matrix_conversions():
    from trackertraincode_b200.eval import PerspectiveCorrector
    from trackertraincode_b200.neuralnets import torchquaternion as tq

    np.testing.assert_allclose(tq.tomatrix(_cuda(G["pose"])).cpu().numpy(), G["tomatrix"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(tq.from_matrix(_cuda(G["tomatrix"])).cpu().numpy(), G["from_matrix"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(tq.from_matrix(_cuda(G["special_mats"])).cpu().numpy(), G["special_from_matrix"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(PerspectiveCorrector._make_look_at_matrix(_cuda(G["look_in"])).cpu().numpy(), G["look_out"], rtol=1e-4, atol=2e-6)
    m = tq.tomatrix(_cuda(G["pose"]).reshape(16, 16, 4))
    assert m.shape == (16, 16, 3, 3)
    with pytest.raises(Exception):
        tq.tomatrix(torch.from_numpy(G["pose"]))  # CPU tensors are refused: no fallback


@pytest.mark.parametrize("fov, image_size, coord, pose, expected", KATS)
def test_reference_perspective_kats(fov, image_size, coord, pose, expected):
    from trackertraincode_b200.eval import PerspectiveCorrector

    q = PerspectiveCorrector(fov).corrected_rotation(torch.as_tensor(image_size, dtype=torch.long), torch.as_tensor(coord, dtype=torch.float32).cuda(),
                                                     torch.from_numpy(pose.as_quat()).to(torch.float32).cuda())
    got = Rotation.from_quat(q.cpu().numpy())
    assert (expected.inv() * got).magnitude() * 180 / math.pi < 0.01
