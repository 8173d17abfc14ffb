"""The deformable face model behind PutRoiFromLandmarks(extend_to_forehead=True) (batch/misc.py:9-31): scaled BFM mean
shape + deformation bases (trackertraincode/facemodel/bfm.py:24-96), posed on the GPU by b200aug_head_roi.

The model data (bfm_noneck_v3.pkl, 24 MB) ships with the reference, not with this package: `HeadModel.from_bfm_pickle(path)`
reads the reference's file; `HeadModel.default()` looks for it in $B200AUG_BFM_PATH or next to an importable
`trackertraincode.facemodel`.  Without it the transform raises -- there is no approximation to fall back to."""
from __future__ import annotations

import ctypes as C
import os
import pickle
from typing import Optional

import numpy as np
import torch

from .. import _native as N

BFM_FILE = "bfm_noneck_v3.pkl"


class HeadModel:
    def __init__(self, vertices, deform_base=None):
        """vertices [V,3] and deform_base [K,V,3] (K <= 64) as ScaledBfmModule holds them (bfm.py:80-96)."""
        self.vertices = torch.as_tensor(vertices, dtype=torch.float32).contiguous()
        self.deform_base = None if deform_base is None else torch.as_tensor(deform_base, dtype=torch.float32).contiguous()
        if self.vertices.dim() != 2 or self.vertices.shape[1] != 3:
            raise ValueError(f"vertices must be [V,3], got {tuple(self.vertices.shape)}")
        if self.deform_base is not None and (self.deform_base.dim() != 3 or tuple(self.deform_base.shape[1:]) != tuple(self.vertices.shape)):
            raise ValueError(f"deform_base must be [K,V,3], got {tuple(self.deform_base.shape)}")
        self._on = {}

    @staticmethod
    def from_bfm_pickle(path: str, shape_dim: int = 40, exp_dim: int = 10) -> "HeadModel":
        """BFMModel.scaled_vertices / scaled_bases (bfm.py:24-72) from the reference's pickle."""
        with open(path, "rb") as f:
            bfm = pickle.load(f)
        u = bfm.get("u").astype(np.float32)
        w_shp = bfm.get("w_shp").astype(np.float32)[..., :shape_dim]
        w_exp = bfm.get("w_exp").astype(np.float32)[..., :exp_dim]
        V = u.shape[0] // 3
        mirror = np.array([[[1.0, -1.0, -1.0]]])
        shp = (20.0 * w_shp.reshape((V, 3, -1))).transpose([2, 0, 1]) * mirror
        exp = (5.0e-5 * w_exp.reshape((V, 3, -1))).transpose([2, 0, 1]) * mirror
        vertices = u.reshape((-1, 3)) * 1.0e-5 * np.array([[1.0, -1.0, -1.0]], dtype="f4")
        vertices = vertices - np.array([0.0, -0.26, -0.9], dtype="f4")[None, :]
        return HeadModel(np.ascontiguousarray(vertices, np.float32), np.concatenate([shp, exp], 0).astype(np.float32))

    @staticmethod
    def default() -> "HeadModel":
        cands = []
        if os.environ.get("B200AUG_BFM_PATH"):
            cands.append(os.environ["B200AUG_BFM_PATH"])
        try:  # an installed / checked-out reference next to this package
            import importlib.util

            spec = importlib.util.find_spec("trackertraincode.facemodel")
            if spec is not None and spec.submodule_search_locations:
                cands += [os.path.join(p, BFM_FILE) for p in spec.submodule_search_locations]
        except (ImportError, ValueError):
            pass
        for p in cands:
            if os.path.isfile(p):
                return HeadModel.from_bfm_pickle(p)
        raise N.NativeError(f"extend_to_forehead needs the reference's face model ({BFM_FILE}): set B200AUG_BFM_PATH or pass a HeadModel")

    def _device_arrays(self, device):
        key = (device.type, device.index)
        hit = self._on.get(key)
        if hit is None:
            hit = (self.vertices.to(device), None if self.deform_base is None else self.deform_base.to(device))
            self._on[key] = hit
        return hit

    def roi(self, coord: torch.Tensor, quat: torch.Tensor, shapeparams: Optional[torch.Tensor] = None, xy_offset: float = 0.0) -> torch.Tensor:
        """[..., 4] = (min_x, min_y, max_x, max_y) over the posed vertices; coord [..., 3], quat [..., 4] CUDA tensors."""
        if not coord.is_cuda:
            raise N.NativeError(f"coord lives on {coord.device}; the B200 path needs CUDA tensors (there is no CPU fallback)")
        prefix = coord.shape[:-1]
        c = coord.reshape(-1, 3).to(torch.float32).contiguous()
        q = quat.reshape(-1, 4).to(coord.device, torch.float32).contiguous()
        B = c.shape[0]
        verts, base = self._device_arrays(coord.device)
        sp, K = None, 0
        if shapeparams is not None:
            if base is None:
                raise ValueError("shape parameters given, but the model has no deformation bases")
            K = base.shape[0]
            sp = shapeparams.reshape(-1, shapeparams.shape[-1]).to(coord.device, torch.float32)
            if sp.shape[-1] != K:
                raise ValueError(f"{sp.shape[-1]} shape parameters for {K} deformation bases")
            sp = sp.expand(B, K).contiguous()
        out = torch.empty((B, 4), dtype=torch.float32, device=coord.device)
        with torch.cuda.device(coord.device):
            N.check(N.lib.b200aug_head_roi(verts.data_ptr(), base.data_ptr() if (base is not None and sp is not None) else None,
                                           verts.shape[0], K, sp.data_ptr() if sp is not None else None, c.data_ptr(), q.data_ptr(),
                                           C.c_float(xy_offset), out.data_ptr(), B,
                                           C.c_void_p(torch.cuda.current_stream(coord.device).cuda_stream)), "b200aug_head_roi")
        out._b200aug_keep = (c, q, sp, verts, base)
        return out.reshape(*prefix, 4)
