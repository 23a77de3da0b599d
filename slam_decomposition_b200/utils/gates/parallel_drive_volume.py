"""Coverage-set Monte-Carlo (reference: src/slam/utils/gates/parallel_drive_volume.py:82-225, 292-307).

The reference draws N = 3000 random parameter vectors per (basis gate, k), evaluates the smush template with
qiskit/qutip, maps every unitary to Weyl coordinates and mirror-folds the cloud before handing it to an exact
convex-hull code (lrs/monodromy, out of scope).  Here the generate -> evaluate -> c1c2c3 -> fold -> bin chain is
one fused CUDA kernel (``slam_coverage_mc``) driven by a counter-based Philox stream, so 1e7..1e9 samples are
practical and any rank can produce exactly its own slice; the hull is replaced by a histogram on the folded
chamber box [0, 1/2]^3 (BASELINE.json configs[1], [4]).
"""
from __future__ import annotations

import numpy as np
import torch

from ... import distributed as D
from ... import engine
from ...basis import CircuitTemplate
from ...basisv2 import CircuitTemplateV2
from .custom_gates import ConversionGainGate, ConversionGainSmushGate

duration_1q = 0.25  # parallel_drive_volume.py:82
N = 3000            # parallel_drive_volume.py:83

# gc, gg, t, name, iters (parallel_drive_volume.py:91-97)
GATE_LIST = [
    (np.pi / 2, 0, 1, "iSwap", 3),
    (np.pi / 2, 0, 1 / 2, "sqiSwap", 3),
    (np.pi / 4, np.pi / 4, 1, "CNOT", 3),
    (np.pi / 4, np.pi / 4, 1 / 2, "sqCNOT", 6),
    (3 * np.pi / 8, np.pi / 8, 1, "B", 2),
    (3 * np.pi / 8, np.pi / 8, 1 / 2, "sqB", 4),
]


def smush_template(gc: float, gg: float, t: float, k: int) -> CircuitTemplateV2:
    """The template of parallel_drive_volume.py:175-199: no exterior 1Q gates, 2Q gate = smush gate whose free
    parameters are (phi_c, phi_g, gx[T], gy[T]) with T = round(t / duration_1q)."""
    T = round(t / duration_1q)

    def pp2(*vargs):
        return ConversionGainSmushGate(vargs[0], vargs[1], gc, gg, vargs[2:2 + T], vargs[2 + T:], t_el=t)

    basis = CircuitTemplateV2(n_qubits=2, base_gates=[pp2], no_exterior_1q=1, vz_only=0, param_vec_expand=[2, T, T])
    basis.build(k)
    return basis


def plain_template(gc: float, gg: float, t: float, k: int) -> CircuitTemplate:
    """Fixed basis gate with interior U3 layers only (exterior 1Q gates do not move Weyl coordinates)."""
    basis = CircuitTemplate(base_gates=[ConversionGainGate(0, 0, gc, gg, t)], no_exterior_1q=(k > 1), maximum_span_guess=k)
    basis.build(k)
    return basis


def _bounds(basis):
    if isinstance(basis, CircuitTemplateV2):
        return basis.default_bound            # basisv2.py:157: every parameter ~ U(-4pi, 4pi)
    return basis.x0_bounds                    # basis.py:111:   U[0, 2pi)


def coverage_points(basis, n_samples: int = N, seed: int = 0, first_sample: int = 0) -> torch.Tensor:
    """Folded, un-rounded Weyl coordinates [n_samples, 3] of random template instances (CUDA tensor)."""
    lo, hi = _bounds(basis)
    _, coords = engine.coverage_mc(basis.desc, seed, first_sample, n_samples, lo, hi, nbins=1, hist=None, want_coords=True)
    return coords


def coverage_histogram(basis, n_samples: int, seed: int = 0, nbins: int = 128, first_sample: int = 0,
                       hist: torch.Tensor | None = None) -> torch.Tensor:
    """int64[nbins^3] histogram of the folded cloud on [0,1/2]^3 (accumulates into `hist` if given)."""
    lo, hi = _bounds(basis)
    hist, _ = engine.coverage_mc(basis.desc, seed, first_sample, n_samples, lo, hi, nbins=nbins, hist=hist)
    return hist


def coverage_sweep(basis, n_samples: int, seed: int = 0, nbins: int = 128) -> torch.Tensor:
    """Multi-GPU form: every rank bins its contiguous shard of the sample stream, then one all-reduce(sum) of the
    histogram.  With a single process this is `coverage_histogram`."""
    rank, world = D.world()
    lo, hi = D.shard_range(n_samples, rank, world)
    hist = coverage_histogram(basis, hi - lo, seed=seed, nbins=nbins, first_sample=lo)
    return D.allreduce_histogram(hist)


def mirror_fold(coordinate_list):
    """parallel_drive_volume.py:295-307: split the cloud into the left (c1 <= 1/2) copy and its mirror image."""
    left, right = [], []
    for x, y, z in coordinate_list:
        if x <= 0.5:
            left.append([x, y, z])
            right.append([1 - x, y, z])
        else:
            left.append([1 - x, y, z])
            right.append([x, y, z])
    return [left, right]


def haar_volume_fraction(hist: torch.Tensor, nbins: int) -> float:
    """Haar-weighted fraction of the Weyl chamber covered by the occupied voxels (voxel analogue of the
    reference's convex-hull volume, parallel_drive_volume.py:343-378).  Density on folded coordinates (units of pi):
    |prod_{i<j} sin(pi(ci+cj)) sin(pi(ci-cj))|."""
    occ = (hist.reshape(nbins, nbins, nbins) > 0)
    g = (torch.arange(nbins, device=hist.device, dtype=torch.float64) + 0.5) / (2 * nbins)
    c1, c2, c3 = torch.meshgrid(g, g, g, indexing="ij")
    inside = (c2 <= c1) & (c3 <= c2)
    pi = np.pi
    dens = (torch.sin(pi * (c1 + c2)) * torch.sin(pi * (c1 - c2)) * torch.sin(pi * (c1 + c3)) * torch.sin(pi * (c1 - c3))
            * torch.sin(pi * (c2 + c3)) * torch.sin(pi * (c2 - c3))).abs()
    w = dens * inside
    return float((w * occ).sum() / w.sum())
