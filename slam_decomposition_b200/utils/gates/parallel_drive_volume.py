"""Coverage-set Monte-Carlo (reference: src/slam/utils/gates/parallel_drive_volume.py:82-225, 292-307).

The reference draws N = 3000 random parameter vectors per (basis gate, k), evaluates the smush template with
qiskit/qutip, maps every unitary to Weyl coordinates and mirror-folds the cloud before handing it to an exact
convex-hull code (lrs/monodromy, out of scope).  Here the generate -> evaluate -> c1c2c3 -> fold -> bin chain is
one fused CUDA kernel (``slam_coverage_mc``) driven by a counter-based Philox stream, so 1e7..1e9 samples are
practical and any rank can produce exactly its own slice; the cloud is binned into a histogram on the folded chamber
box [0, 1/2]^3 (BASELINE.json configs[1], [4]).  The post-processing the reference does on the cloud -- convex hull per
mirror side, union with the base polytope, Haar volume, CNOT / SWAP / B membership (parallel_drive_volume.py:343-396) --
is `hull_coverage` below: a host-side step on the occupied voxels (it is not on the hot path).
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


# ----------------------------------------------------------------------------------------------------------------------
# Post-processing: hull / Haar volume / membership flags (reference: parallel_drive_volume.py:292-396, results in
# src/slam/data/extended_results.json as [base_vol, extended_vol, has_CNOT, has_SWAP, has_B] per k).
#
# The reference mirror-folds the cloud into a left (c1 <= 1/2) and a right copy, takes the convex hull of each side with
# lrs, unions both hulls with the analytic polytope of k plain basis gates (monodromy) and integrates the Haar measure over
# the union.  Both sides are mirror images and the Haar density is symmetric under c1 -> 1 - c1, so the volume fraction of
# the union equals the fraction of the LEFT half-chamber covered by (left hull) U (base set) -- which is what the folded
# histogram holds.  monodromy's coordinates are an affine image of (c1, c2, c3), so the hull may be taken in either.
# The base polytope needs monodromy's exact-rational geometry (out of scope); here the base set is the convex hull of the
# PLAIN template's cloud (k fixed basis gates with free interior 1Q gates), which reproduces the reference's base volumes
# wherever that set is convex.
# ----------------------------------------------------------------------------------------------------------------------
WEYL_POINTS = {"CNOT": (0.5, 0.0, 0.0), "SWAP": (0.5, 0.5, 0.5), "B": (0.5, 0.25, 0.0)}


def haar_density(c: np.ndarray) -> np.ndarray:
    """Un-normalised Haar density on Weyl coordinates in units of pi: |prod_{i<j} sin(pi(ci+cj)) sin(pi(ci-cj))|."""
    c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2]
    pi = np.pi
    return np.abs(np.sin(pi * (c1 + c2)) * np.sin(pi * (c1 - c2)) * np.sin(pi * (c1 + c3)) * np.sin(pi * (c1 - c3))
                  * np.sin(pi * (c2 + c3)) * np.sin(pi * (c2 - c3)))


def half_chamber_grid(n: int = 256):
    """Midpoint grid (n^3 cells on [0, 1/2]^3) restricted to the left half-chamber 1/2 >= c1 >= c2 >= c3 >= 0:
    (points [M, 3], Haar weights [M])."""
    g = (np.arange(n) + 0.5) / (2 * n)
    c1, c2, c3 = np.meshgrid(g, g, g, indexing="ij")
    inside = (c2 <= c1) & (c3 <= c2)
    pts = np.stack([c1[inside], c2[inside], c3[inside]], axis=1)
    return pts, haar_density(pts)


def occupied_voxel_centres(hist, nbins: int) -> np.ndarray:
    """Centres [M, 3] of the occupied voxels of a folded coverage histogram (torch tensor or numpy array)."""
    h = hist.detach().cpu().numpy() if isinstance(hist, torch.Tensor) else np.asarray(hist)
    idx = np.argwhere(h.reshape(nbins, nbins, nbins) > 0)
    return (idx + 0.5) / (2 * nbins)


def hull_points(hist, nbins: int) -> np.ndarray:
    """Points whose convex hull stands for the cloud: the centre of every occupied voxel -- for a densely sampled set the
    last occupied voxel along a direction is cut by the boundary at a uniformly distributed position, so the centre is the
    unbiased choice -- plus, for voxels that touch a face of the half-chamber (c3 = 0, c1 = 1/2, c1 = c2, c2 = c3), their
    corners clamped into the chamber: the set reaches those faces exactly, and the Haar density does not vanish on c3 = 0
    and c1 = 1/2, so a half-voxel gap there would cost volume."""
    h = hist.detach().cpu().numpy() if isinstance(hist, torch.Tensor) else np.asarray(hist)
    idx = np.argwhere(h.reshape(nbins, nbins, nbins) > 0)
    if not len(idx):
        return np.zeros((0, 3))
    step = 1.0 / (2 * nbins)
    pts = [(idx + 0.5) * step]
    touch = (idx[:, 2] == 0) | (idx[:, 0] == nbins - 1) | (idx[:, 0] == idx[:, 1]) | (idx[:, 1] == idx[:, 2])
    edge = idx[touch]
    for d in np.ndindex(2, 2, 2):
        c = (edge + np.array(d)) * step
        c[:, 0] = np.minimum(c[:, 0], 0.5)
        c[:, 1] = np.minimum(c[:, 1], c[:, 0])
        c[:, 2] = np.minimum(c[:, 2], c[:, 1])
        pts.append(c)
    return np.concatenate(pts)


def _cloud_points(cloud, nbins: int) -> np.ndarray:
    """A cloud is either folded coordinates [M, 3] (exact sample points, e.g. `coverage_points`) or a folded histogram
    int64[nbins^3] (voxel approximation: `hull_points`; the most protruding voxel centre lies up to sqrt(3)/2 voxel
    outside an oblique face, so prefer points where the sample count allows)."""
    a = cloud.detach().cpu().numpy() if isinstance(cloud, torch.Tensor) else np.asarray(cloud)
    if a.ndim == 2 and a.shape[1] == 3 and a.dtype.kind == "f":
        return a
    return hull_points(a, nbins)


def _hull_equations(points: np.ndarray):
    """Facet inequalities (A, b) with A x + b <= 0 inside, or None for fewer than 4 points / a flat cloud (zero volume)."""
    import scipy.spatial as sp

    if points.shape[0] < 4 or np.linalg.matrix_rank(points - points[0], tol=1e-9) < 3:
        return None
    try:
        h = sp.ConvexHull(points)
    except sp.QhullError:
        return None
    return h.equations[:, :3].copy(), h.equations[:, 3].copy()


def _inside(eq, pts: np.ndarray, tol: float = 1e-12) -> np.ndarray:
    if eq is None:
        return np.zeros(len(pts), dtype=bool)
    A, b = eq
    out = np.empty(len(pts), dtype=bool)
    step = max(1, 4_000_000 // max(1, len(b)))
    for i in range(0, len(pts), step):
        out[i:i + step] = (pts[i:i + step] @ A.T + b).max(axis=1) <= tol
    return out


def hull_coverage(smush_cloud, plain_cloud=None, nbins: int = 128, grid: int = 256, base_flags=None) -> dict:
    """[base_vol, extended_vol, has_CNOT, has_SWAP, has_B] of parallel_drive_volume.py:343-396 from folded clouds:
    `smush_cloud` of the parallel-drive template, `plain_cloud` of the plain template with the same k (the base set; None =
    empty); each either sample coordinates [M, 3] or a histogram int64[nbins^3].  Volumes are Haar fractions of the Weyl
    chamber.  Membership flags: the point lies inside the cloud's hull, or the base set contains it -- `base_flags` =
    {"CNOT": bool, ...} from `base_reachable` (an exact decomposition test on the plain template), else a point of the plain
    cloud within one voxel of it."""
    pts, w = half_chamber_grid(grid)
    ps = _cloud_points(smush_cloud, nbins)
    pb = _cloud_points(plain_cloud, nbins) if plain_cloud is not None else None
    eq_s = _hull_equations(ps)
    eq_b = _hull_equations(pb) if pb is not None else None
    in_b = _inside(eq_b, pts)
    in_s = _inside(eq_s, pts)
    total = w.sum()
    out = {"base_vol": float((w * in_b).sum() / total), "extended_vol": float((w * (in_b | in_s)).sum() / total)}
    voxel = 1.0 / (2 * nbins)
    for name, p in WEYL_POINTS.items():
        q = np.array([p])
        if base_flags is not None:
            in_base = bool(base_flags[name])
        elif pb is not None:
            in_base = bool(len(pb)) and bool(np.abs(pb - q).max(axis=1).min() <= voxel)  # a cloud point within one voxel
        else:
            in_base = False
        out[f"has_{name}"] = bool(in_base or _inside(eq_s, q, tol=0.0)[0])
    return out


def base_reachable(gc: float, gg: float, t: float, k: int, restarts: int = 32, threshold: float = 1e-8) -> dict:
    """Exact form of `circuit_polytope.has_element(target)` for the base polytope (parallel_drive_volume.py:381-396): is
    CNOT / SWAP / B reachable by k plain basis gates with free 1Q gates?  Decided by decomposing the three gates onto the
    plain template at this k on the device (loss <= threshold)."""
    from ...cost_function import BasicCost
    from ...optimizer import TemplateOptimizer
    from .custom_gates import BerkeleyGate, CXGate, SwapGate

    basis = CircuitTemplate(base_gates=[ConversionGainGate(0, 0, gc, gg, t)], maximum_span_guess=k, preseed=False)
    opt = TemplateOptimizer(basis=basis, objective=BasicCost(), override_fail=True, training_restarts=restarts,
                            success_threshold=threshold)
    V = np.stack([np.asarray(g.to_matrix() if hasattr(g, "to_matrix") else np.array(g), dtype=np.complex128)
                  for g in (CXGate(), SwapGate(), BerkeleyGate())])
    res = opt.approximate_targets(V, range(k, k + 1), reuse_host_buffers=False)
    return {name: bool(res["success"][i]) for i, name in enumerate(("CNOT", "SWAP", "B"))}


def coverage_study(gc: float, gg: float, t: float, k: int, n_samples: int = N, seed: int = 0, n_base: int = 2_000_000,
                   grid: int = 128, exact_flags: bool = True) -> list:
    """One (gate, k) row of the reference's study (parallel_drive_volume.py:140-410) in its own protocol:
    `n_samples` random instances of the parallel-drive template (reference: N = 3000) -> folded cloud -> hull; base set =
    hull of `n_base` instances of the plain template (the reference takes the exact monodromy polytope, which the plain
    cloud converges to); -> [base_vol, extended_vol, has_CNOT, has_SWAP, has_B] as stored in data/extended_results.json."""
    smush = coverage_points(smush_template(gc, gg, t, k), n_samples, seed=seed)
    plain = coverage_points(plain_template(gc, gg, t, k), n_base, seed=seed + 1)
    flags = base_reachable(gc, gg, t, k) if exact_flags else None
    out = hull_coverage(smush, plain, grid=grid, base_flags=flags)
    return [out["base_vol"], out["extended_vol"], out["has_CNOT"], out["has_SWAP"], out["has_B"]]
