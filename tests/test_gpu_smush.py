"""K4 parity: smush templates (per-slice exp(-i dt H)), constant-gate lowering, gate __array__, trajectory."""
import numpy as np
import pytest
import torch

import oracle as O
from helpers import BASES, make_pair
from slam_decomposition_b200 import engine
from slam_decomposition_b200.hamiltonian import ConversionGainPhaseHamiltonian, ConversionGainSmush, ConversionGainSmush1QPhase
from slam_decomposition_b200.utils.gates.custom_gates import (
    BerkeleyGate, CanonicalGate, ConversionGainGate, ConversionGainSmush1QPhaseGate, ConversionGainSmushGate, RiSwapGate)

pytestmark = pytest.mark.gpu
TOL_U = 1e-10


def _dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


@pytest.mark.parametrize("base,T", [("sqiSwap", 2), ("iSwap", 4), ("sqCNOT", 2), ("B", 4)])
@pytest.mark.parametrize("k", [1, 2, 3])
def test_mc_smush_template_matches_oracle(base, T, k):
    """The coverage template of parallel_drive_volume.py:175-198: no exterior 1Q, Q = (phi_c, phi_g, gx[T], gy[T])."""
    gc, gg, t = BASES[base]
    slots = ("Q", "Q", gc, gg) + ("Q",) * (2 * T) + (t,)
    desc, orc = make_pair("smush", slots, k=k, T=T, no_exterior_1q=True)
    assert orc.n_params == 6 * (k - 1) + k * (2 + 2 * T)
    rng = np.random.default_rng(k)
    X = rng.uniform(-4 * np.pi, 4 * np.pi, (40, orc.n_params))
    U = engine.template_eval(desc, _dev(X)).cpu().numpy()
    Uo = orc.eval_batch(X)
    assert np.abs(U - Uo).max() < TOL_U
    assert np.abs(np.einsum("bij,bkj->bik", U, U.conj()) - np.eye(4)).max() < 1e-12


def test_smush1q_template_and_partial_binding():
    T = 3
    slots = (0.3, "Q", -0.2, 0.5, np.pi / 2, 0.1, "Q", 0.4) + ("Q",) * T + (0.7, "Q", -1.1) + ("Q",)
    desc, orc = make_pair("smush1q", slots, k=2, T=T)
    rng = np.random.default_rng(5)
    X = rng.uniform(-3, 3, (25, orc.n_params))
    U = engine.template_eval(desc, _dev(X)).cpu().numpy()
    assert np.abs(U - orc.eval_batch(X)).max() < TOL_U


def test_constant_smush_gate_is_lowered_on_device_and_differentiable():
    """All-constant smush gates become dense matrices computed by the device; loss/grad then run the adjoint."""
    T = 2
    slots = (0.2, -0.3, np.pi / 2, 0.0, 1.0, -2.0, 0.5, 0.25, 0.5)
    desc, orc = make_pair("smush", slots, k=2, T=T)
    rng = np.random.default_rng(6)
    X = rng.uniform(0, 2 * np.pi, (10, orc.n_params))
    V = O.haar_unitary(rng, 3)
    U = engine.template_eval(desc, _dev(X)).cpu().numpy()
    assert np.abs(U - orc.eval_batch(X)).max() < TOL_U
    loss, grad, _ = engine.loss_grad(desc, _dev(X), _dev(V))
    for b in range(10):
        l, g, _ = O.loss_and_grad(orc, X[b], V[b % 3])
        assert abs(loss[b].item() - l) < 1e-12 and np.abs(grad[b].cpu().numpy() - g).max() < 1e-10


def test_large_drive_amplitudes_need_many_squarings():
    slots = ("Q", "Q", np.pi / 2, 0.0, "Q", "Q", 1.0)
    desc, orc = make_pair("smush", slots, k=1, T=1, no_exterior_1q=True)
    X = np.array([[0.1, -0.2, 200.0, -150.0], [3.0, 1.0, 1e-9, 0.0], [0.0, 0.0, 0.0, 0.0]])
    U = engine.template_eval(desc, _dev(X)).cpu().numpy()
    assert np.abs(U - orc.eval_batch(X)).max() < 1e-9  # ||dt H|| ~ 350: error grows with the 11 squarings


def test_gate_objects_give_device_matrices():
    assert np.abs(np.asarray(RiSwapGate(0.5)) - O.riswap(0.5)).max() < 1e-15
    assert np.abs(np.asarray(ConversionGainGate(0.3, -0.2, 1.1, 0.4, 0.7)) - O.conversion_gain(0.3, -0.2, 1.1, 0.4, 0.7)).max() < 1e-15
    g = ConversionGainSmushGate(0.1, 0.2, 1.0, 0.5, [0.3, -0.4], [1.0, 2.0], 0.5)
    assert np.abs(np.asarray(g) - O.smush(0.1, 0.2, 1.0, 0.5, [0.3, -0.4], [1.0, 2.0], 0.5)).max() < TOL_U
    g1 = ConversionGainSmush1QPhaseGate(0.1, 0.2, 0.3, 0.4, 1.0, 0.5, 0.3, -0.2, [1.0, -2.0, 0.5], [0.3, 0.1, -1.0], 1.5)
    ref = O.smush_1qphase(0.1, 0.2, 0.3, 0.4, 1.0, 0.5, 0.3, -0.2, [1.0, -2.0, 0.5], [0.3, 0.1, -1.0], 1.5)
    assert np.abs(np.asarray(g1) - ref).max() < TOL_U
    assert np.abs(np.asarray(BerkeleyGate()) - O.berkeley()).max() < 1e-15
    assert np.abs(np.asarray(CanonicalGate(0.3, 0.2, 0.1)) - O.canonical_gate(0.6 / np.pi, 0.4 / np.pi, 0.2 / np.pi)).max() < 1e-15
    assert g.cost() == pytest.approx(1.5 * 0.5 / (np.pi / 2))


def test_hamiltonian_factories_keep_reference_conventions():
    # positional quirk of ConversionGainPhaseHamiltonian.construct_U
    U = ConversionGainPhaseHamiltonian.construct_U(0.3, -0.2, 1.1, 0.4, t=0.7)
    assert np.abs(U - O.conversion_gain(0.3, -0.2, 1.1, 0.4, 0.7)).max() < 1e-15
    U = ConversionGainSmush.construct_U(0.1, 0.2, 1.0, 0.5, [0.3, -0.4], [1.0, 2.0], t=0.5)
    assert np.abs(U - O.smush(0.1, 0.2, 1.0, 0.5, [0.3, -0.4], [1.0, 2.0], 0.5)).max() < TOL_U
    U = ConversionGainSmush1QPhase.construct_U(0.1, 0.2, 0.3, 0.4, 1.0, 0.5, 0.3, -0.2, [1.0], [0.3], t=0.1)
    assert np.abs(U - O.smush_1qphase(0.1, 0.2, 0.3, 0.4, 1.0, 0.5, 0.3, -0.2, [1.0], [0.3], 0.1)).max() < TOL_U


def test_trajectory_matches_iterate_time():
    rng = np.random.default_rng(9)
    B, N, R, dt = 6, 10, 5, 0.1
    gate = np.zeros((B, 8))
    gate[:, :4] = rng.uniform(-1, 1, (B, 4))
    gate[:, 4] = np.pi / 2
    gate[:, 5] = rng.uniform(0, 0.5, B)
    gate[:, 6:] = rng.uniform(-0.5, 0.5, (B, 2))
    gate[0, :4] = 0.0
    gate[0, 5:] = 0.0  # the widget default: iSWAP drive, no phases
    gx = rng.uniform(-2 * np.pi, 2 * np.pi, (B, N))
    gy = rng.uniform(-2 * np.pi, 2 * np.pi, (B, N))
    coords, Uf = engine.pd_trajectory(_dev(gate), _dev(gx), _dev(gy), dt, R=R, fold=True, round8=False)
    coords, Uf = coords.cpu().numpy(), Uf.cpu().numpy()
    for b in range(B):
        ref_c, ref_U = O.trajectory(gate[b, :4], gate[b, 4], gate[b, 5], gate[b, 6], gate[b, 7], gx[b], gy[b], dt, R)
        assert np.abs(Uf[b] - ref_U).max() < TOL_U
        # the oracle rounds to 8 dp like the reference; compare the un-rounded device coordinates against it
        assert np.abs(coords[b] - ref_c).max() < 1e-8
    # rounded + folded output equals the reference's stored values except at rounding ties
    c8, _ = engine.pd_trajectory(_dev(gate), _dev(gx), _dev(gy), dt, R=R, fold=True, round8=True)
    ref_all = np.stack([O.trajectory(gate[b, :4], gate[b, 4], gate[b, 5], gate[b, 6], gate[b, 7], gx[b], gy[b], dt, R)[0]
                        for b in range(B)])
    assert np.mean(np.abs(c8.cpu().numpy() - ref_all) < 1e-12) > 0.97
    assert np.abs(c8.cpu().numpy() - ref_all).max() <= 1.0000001e-8


def test_parallel_driven_gate_widget_matches_reference_semantics():
    from slam_decomposition_b200.utils.pd_playground import ParallelDrivenGateWidget, trajectories

    w = ParallelDrivenGateWidget(N=10, gc=np.pi / 2, gg=0)
    w.prepare_parameters_nonuniform([3] * w.N, [0] * w.N)  # ImprovedCX (pd_playground.py:255-258)
    w.iterate_time(R=5)
    ref_c, ref_U = O.trajectory((0, 0, 0, 0), np.pi / 2, 0, 0, 0, [3] * 10, [0] * 10, 0.1, 5)
    got = np.array(w.coordinate_list)
    assert got.shape == (10, 5, 3) and np.abs(got - ref_c).max() <= 1.0000001e-8
    assert np.abs(w.final_unitary - ref_U).max() < TOL_U
    assert np.abs(w.solve_end() - ref_U).max() < TOL_U
    assert w.end_segment_list[-1] == w.coordinate_list[-1][-1]
    # default widget (no 1Q drive): after 10 x 0.1 of an iSWAP-strength drive the gate is iSWAP
    w0 = ParallelDrivenGateWidget()
    assert np.allclose(O.c1c2c3(w0.solve_end()), (0.5, 0.5, 0.0), atol=1e-8)
    c, U = trajectories(np.zeros((3, 8)), np.ones((3, 4)), np.ones((3, 4)))
    assert c.shape == (3, 4, 5, 3) and U.shape == (3, 4, 4)


def test_drive_amplitude_search_with_makhlin_cost():
    """scripts/parallel_drive_swap (cells 7-10): search the per-slice drive amplitudes of a parallel-driven iSWAP pulse
    (N = 10 slices of 0.1) for a target class with a Makhlin cost and Nelder-Mead.  The N-slice widget circuit is one
    ConversionGainSmush1QPhaseGate with T = N, so TemplateOptimizer drives it through the generic Nelder-Mead kernel."""
    from slam_decomposition_b200.basisv2 import CircuitTemplateV2
    from slam_decomposition_b200.cost_function import MakhlinEuclideanCost
    from slam_decomposition_b200.optimizer import TemplateOptimizer
    from slam_decomposition_b200.utils.pd_playground import ParallelDrivenGateWidget

    N, dt = 10, 0.1

    def pulse(*v):
        return ConversionGainSmush1QPhaseGate(0, 0, 0, 0, np.pi / 2, 0, 0, 0, v[:N], v[N:], t_el=N * dt)

    basis = CircuitTemplateV2(n_qubits=2, base_gates=[pulse], no_exterior_1q=True, param_vec_expand=[0, N, N])
    basis.build(1)
    basis.spanning_range = range(1, 2)
    assert basis.desc.n_params == 2 * N and basis.desc.T == N
    # the template with amplitudes x equals the widget's solve_end() with the same non-uniform drive
    x = np.linspace(-3, 3, 2 * N)
    names = [p.name for p in basis.circuit.parameters]
    natural = np.array([x[int(n[1:])] for n in names])  # Xk is name-sorted (Q0, Q1, Q10, ...)
    w = ParallelDrivenGateWidget(N=N, gc=np.pi / 2)
    w.prepare_parameters_nonuniform(x[:N], x[N:])
    assert np.abs(basis.eval(natural) - w.solve_end()).max() < 1e-10
    # search: reach the CNOT class (the "ImprovedCX" use case, pd_playground.py:251-258) from random amplitudes
    np.random.seed(12)
    for b in names:
        basis.bounds[b] = (-4.0, 4.0)  # initial-point box only (no add_bound -> unbounded optimisation)
    opt = TemplateOptimizer(basis, MakhlinEuclideanCost(), override_fail=True, training_restarts=64, success_threshold=1e-7)
    d = opt.approximate_target_U(O.CNOT)
    U = basis.eval(d.Xk)
    assert abs(O.cost(U, O.CNOT, "makhlin_euclidean") - d.loss_result) < 3e-8
    assert d.loss_result < 5e-3 and np.allclose(O.fold_c1(np.array(O.c1c2c3(U))), (0.5, 0.0, 0.0), atol=5e-3)


def test_b11_reference_recorded_smush_solution_on_device():
    """KAT B11 (scripts/local_smush_test.ipynb cell 5): the reference's own solved smush circuit, T = 5 slices between
    exterior U3 layers, through K1 (unitary vs oracle at 1e-10), K3 (coordinates), the fused cost functional and K5b's
    objective (MakhlinFunctionalCost vs CX == 0.0 exactly, as the reference logged 'Best Loss=0.0')."""
    import json
    import os

    from helpers import b11_case
    from slam_decomposition_b200 import _lib
    from slam_decomposition_b200.cost_function import MakhlinFunctionalCost

    kats = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kats.json")))
    slots, k, T, vals = b11_case(kats)
    desc, orc = make_pair("smush", slots, k=k, T=T)
    x = np.array([[vals[n] for n in orc.names_sorted]])
    U = engine.template_eval(desc, _dev(x))
    assert np.abs(U.cpu().numpy()[0] - orc.eval(x[0])).max() < TOL_U
    c, g = engine.weyl(U, round8=True, want_g=True)
    c = c[0].tolist()
    assert c[0] == kats["B11"]["c1"] and 0 <= c[1] < kats["B11"]["c2_c3_below"] and 0 <= c[2] < kats["B11"]["c2_c3_below"]
    assert tuple(c) == O.c1c2c3(orc.eval(x[0]))
    assert tuple(abs(v) for v in g[0].tolist()) == (0.0, 0.0, 1.0)
    assert MakhlinFunctionalCost().unitary_fidelity(U[0].cpu().numpy(), O.CNOT) == kats["B11"]["makhlin_functional_vs_CX"] == 0.0
    # K5b objective: Nelder-Mead started AT the recorded solution can only report a loss <= f(x0) = 0.0, and the
    # functional is non-negative, so the kernel's own evaluation of the recorded point must be exactly 0.0
    nm = engine.nm_defaults()
    nm.cost_kind, nm.max_iter, nm.early_exit = _lib.COST_MAKHLIN_FUNCTIONAL, 1, 0
    loss, _, _ = engine.nm_solve(desc, _dev(O.CNOT[None].astype(np.complex128)), 1, nm, x0=_dev(x[None]))
    assert loss.item() == 0.0
