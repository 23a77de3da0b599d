"""Target distributions (reference: src/slam/sampler.py).  In scope: ``HaarSample`` and ``GateSample``;
``CircuitSample`` / ``Clifford`` / ``Haar2Sample`` / ``Haar3Sample`` need qiskit's transpiler.

``HaarSample`` draws exactly the stream of the reference (python ``random`` -> seed of
``scipy.stats.unitary_group``, which is what qiskit's ``random_unitary`` calls), including the quirk
that it re-seeds on every call.  ``batch(device)`` hands all targets to the GPU at once.
"""
from __future__ import annotations

import random
from abc import ABC
from sys import maxsize

import numpy as np
import scipy.stats


class SampleFunction(ABC):
    def __init__(self, n_qubits=2, n_samples=1):
        self.n_qubits = n_qubits
        self.n_samples = n_samples

    def __iter__(self):
        for _ in range(self.n_samples):
            yield self._get_unitary()

    def _get_unitary(self):
        raise NotImplementedError

    def batch(self, device=None):
        """All targets as one [n_samples, 4, 4] complex128 CUDA tensor."""
        import torch

        from . import engine

        device = device or engine.require_cuda()
        arr = np.stack([np.asarray(u, dtype=np.complex128) for u in self])
        return torch.as_tensor(arr, device=device)


class GateSample(SampleFunction):
    def __init__(self, gate, n_samples=1):
        self.gate = gate
        super().__init__(gate.num_qubits, n_samples)

    def _get_unitary(self):
        return np.asarray(self.gate.__array__(), dtype=np.complex128)


class HaarSample(SampleFunction):
    def __init__(self, seed=None, n_samples=1, n_qubits=2):
        self.seed = seed
        super().__init__(n_samples=n_samples, n_qubits=n_qubits)

    def _get_unitary(self):
        random.seed(self.seed)  # sampler.py:67-71: re-seeded on every call
        s = random.randint(0, maxsize)
        return scipy.stats.unitary_group.rvs(2 ** self.n_qubits, random_state=np.random.default_rng(s))


class HaarBatchSample(SampleFunction):
    """Extension for large sweeps: `n_samples` independent Haar targets from one seeded generator
    (the reference's HaarSample with a fixed seed yields the same matrix every time)."""

    def __init__(self, seed=0, n_samples=1):
        self.seed = seed
        super().__init__(n_samples=n_samples, n_qubits=2)

    def __iter__(self):
        rng = np.random.default_rng(self.seed)
        for _ in range(self.n_samples):
            yield scipy.stats.unitary_group.rvs(4, random_state=rng)
