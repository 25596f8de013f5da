"""A minimal stand-in for spatialmath's SE3 (spatialmath-python is not a dependency here).

``ETS.fkine`` in the reference wraps every pose of a trajectory into its own SE3 object
(reference ETS.py:1013-1015) -- O(N) Python objects, which for a 1M-row batch costs far more
than the kinematics.  This container keeps the batch as one (N,4,4) array and only exposes the
few accessors the hot path's callers use (.A, .t, .R, len, indexing, iteration)."""
from __future__ import annotations

import numpy as np


class SE3:
    def __init__(self, A):
        A = getattr(A, "A", A)
        A = np.asarray(A)
        if A.shape[-2:] != (4, 4):
            raise ValueError("SE3 needs (4,4) or (N,4,4) data")
        self._A = A.reshape(-1, 4, 4)
        self._single = A.ndim == 2

    @property
    def A(self):
        return self._A[0] if self._single else self._A

    @property
    def t(self):
        return self.A[..., :3, 3]

    @property
    def R(self):
        return self.A[..., :3, :3]

    def __len__(self):
        return self._A.shape[0]

    def __getitem__(self, i):
        return SE3(self._A[i])

    def __iter__(self):
        for i in range(self._A.shape[0]):
            yield SE3(self._A[i])

    def __repr__(self):
        return f"SE3(n={len(self)})\n{self.A}"
