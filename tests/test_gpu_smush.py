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


def _segments_case(rng, B, sizes):
    """Random multi-segment pulses: per segment its own phases / couplings, per slice its own amplitudes."""
    N = sum(sizes)
    gate = np.zeros((B, N, 8))
    for b in range(B):
        s0 = 0
        for n in sizes:
            gate[b, s0:s0 + n] = np.concatenate([rng.uniform(-np.pi, np.pi, 4), rng.uniform(-2, 2, 4)])
            s0 += n
    return gate, rng.uniform(-2 * np.pi, 2 * np.pi, (B, N)), rng.uniform(-2 * np.pi, 2 * np.pi, (B, N))


def test_multi_segment_trajectory_matches_the_composed_widgets():
    """`pdgw + pdgw2 + pdgw3` (pd_playground.py:46-58): per-slice gate rows through slam_pd_trajectory_slices against the
    oracle's restatement of the composed circuit + iterate_time -- final unitary and un-rounded coordinates at 1e-10."""
    rng = np.random.default_rng(17)
    B, sizes, dt, R = 40, (10, 3, 2), 0.1, 5
    gate, gx, gy = _segments_case(rng, B, sizes)
    coords, Uf = engine.pd_trajectory(_dev(gate), _dev(gx), _dev(gy), dt, R=R, fold=True, round8=False)
    c8, _ = engine.pd_trajectory(_dev(gate), _dev(gx), _dev(gy), dt, R=R, fold=True, round8=True)
    for b in range(B):
        ref_c, ref_U = O.trajectory_segments(gate[b], gx[b], gy[b], dt, R)
        assert np.abs(Uf[b].cpu().numpy() - ref_U).max() < TOL_U
        assert np.abs(coords[b].cpu().numpy() - ref_c).max() < 1e-8 + 1e-10  # the oracle's coordinates are 8-dp rounded
        assert np.abs(c8[b].cpu().numpy() - ref_c).max() <= 1e-8 + 1e-12
    # one gate row per slice with identical rows == the single-row entry point, bit for bit
    same = np.repeat(gate[:, :1], sum(sizes), axis=1)
    a = engine.pd_trajectory(_dev(same), _dev(gx), _dev(gy), dt, R=R)[1]
    b_ = engine.pd_trajectory(_dev(same[:, 0]), _dev(gx), _dev(gy), dt, R=R)[1]
    assert torch.equal(a, b_)
    with pytest.raises(ValueError):
        engine.pd_trajectory(_dev(gate[:, :5]), _dev(gx), _dev(gy), dt)


def test_widget_addition_improved_cx_and_swap():
    """ParallelDrivenGateWidget.__add__, ImprovedCX, ImprovedSWAP (pd_playground.py:46-58, 247-339) against the oracle."""
    from slam_decomposition_b200.utils.pd_playground import ImprovedCX, ImprovedSWAP, ParallelDrivenGateWidget

    # notebook cell 6: pdgw (N = 10) + pdgw2 (N = 5, phase_a = pi/2, phase_b = -pi/8)
    w = ParallelDrivenGateWidget(N=10, gc=np.pi / 2, gg=0) + ParallelDrivenGateWidget(N=5, gc=np.pi / 2, gg=0, phase_a=np.pi / 2,
                                                                                       phase_b=-np.pi / 8)
    assert w.N == 15
    gx = list(np.ones(10) * np.pi) + list(np.ones(5) * 2 * np.pi)
    gy = list(np.ones(10) * np.pi) + list(np.ones(5) * np.pi / 2)
    w.prepare_parameters_nonuniform(gx, gy)
    w.iterate_time()
    rows = np.zeros((15, 8))
    rows[:, 4] = np.pi / 2
    rows[10:, 0], rows[10:, 1] = np.pi / 2, -np.pi / 8
    ref_c, ref_U = O.trajectory_segments(rows, gx, gy, 0.1, 5)
    assert np.abs(np.array(w.coordinate_list) - ref_c).max() <= 1e-8 + 1e-12
    assert np.abs(w.final_unitary - ref_U).max() < TOL_U and np.abs(w.solve_end() - ref_U).max() < TOL_U
    three = w + ParallelDrivenGateWidget(N=2, gc=np.pi / 2, phase_c=0.3)
    assert three.N == 17 and three._gate_rows.shape == (17, 8) and three._gate_rows[16, 2] == 0.3 and three._gate_rows[12, 0] == np.pi / 2
    # ImprovedCX: gx = 3 on the ten slices of the default iSWAP-strength drive
    cx = ImprovedCX()
    ref_c, _ = O.trajectory((0, 0, 0, 0), np.pi / 2, 0, 0, 0, [3] * 10, [0] * 10, 0.1, 5)
    assert np.abs(np.array(cx.coordinate_list[:10]) - ref_c).max() <= 1e-8 + 1e-12
    assert cx.coordinate_list[10] == [[0, 0, 0]] * 5 and cx.coordinate_list[11] == [(0.5, 0, 0)] * 5
    assert len(cx.baseline_coords) == 5 and cx.baseline_coords[1][-1] == [0.5, 0.0, 0]
    # ImprovedSWAP: gx = gy = pi, then the two recorded U3 gates on qubit 0
    sw = ImprovedSWAP()
    _, ref_U = O.trajectory((0, 0, 0, 0), np.pi / 2, 0, 0, 0, [np.pi] * 10, [np.pi] * 10, 0.1, 5)
    for tri in ImprovedSWAP.TAIL_U3:
        ref_U = np.kron(np.eye(2), O.u3(*tri)) @ ref_U
    cref = list(O.c1c2c3(ref_U))
    cref[0] = 1 - cref[0] if cref[0] > 0.5 else cref[0]
    assert np.abs(sw.extended_unitary - ref_U).max() < TOL_U
    assert len(sw.coordinate_list) == 10 + 50 + 2 and np.abs(np.array(sw.coordinate_list[10]) - np.array(cref)).max() <= 1e-8 + 1e-12


def test_batched_pulse_search_reproduces_the_swap_objective_of_the_notebook():
    """scripts/parallel_drive_swap/parallel_drive_swap.ipynb cell 7: ten fixed slices (gx = gy = pi) + a five-slice tail with
    four free phases and two amplitudes TIED across the tail's slices, MakhlinFunctionalCost against SWAP, Nelder-Mead --
    here 4096 starts at once in K5b on a 15-gate template with tied slots.  (The notebook's own cost_function evaluates
    `pdgw.solve_end()` instead of `pdgw3.solve_end()`, so its logged 'function value 5.0' is the constant cost of the fixed
    prefix; the intended objective is restated.)"""
    from slam_decomposition_b200.circuit import Parameter
    from slam_decomposition_b200.cost_function import MakhlinFunctionalCost
    from slam_decomposition_b200.utils.pd_playground import pulse_template, search_pulse

    p = [Parameter(f"A{i}") for i in range(6)]
    segs = [dict(N=10, gc=np.pi / 2, gx=np.pi, gy=np.pi),
            dict(N=5, gc=np.pi / 2, phase_a=p[0], phase_b=p[1], phase_c=p[2], phase_g=p[3], gx=p[4], gy=p[5])]
    desc, names = pulse_template(segs)
    assert (desc.k, desc.T, desc.n_params, names) == (15, 1, 6, [f"A{i}" for i in range(6)])

    def rows(x):
        r = np.zeros((15, 8))
        r[:, 4] = np.pi / 2
        r[10:, :4] = x[:4]
        return r, [np.pi] * 10 + [x[4]] * 5, [np.pi] * 10 + [x[5]] * 5

    def oracle_cost(x):
        r, gx, gy = rows(x)
        return O.J_T_LI(O.SWAP, O.trajectory_segments(r, gx, gy, 0.1, 2)[1])

    # the tied template evaluates to the composed pulse
    rng = np.random.default_rng(4)
    X = rng.uniform(-2 * np.pi, 2 * np.pi, (16, 6))
    U = engine.template_eval(desc, _dev(X)).cpu().numpy()
    for b in range(16):
        r, gx, gy = rows(X[b])
        assert np.abs(U[b] - O.trajectory_segments(r, gx, gy, 0.1, 2)[1]).max() < TOL_U
    # analytic-gradient entry points refuse tied templates instead of producing a wrong gradient
    with pytest.raises(NotImplementedError):
        engine.loss_grad(desc, _dev(X), _dev(O.SWAP[None].astype(np.complex128)))
    # the value the notebook logged: the undriven prefix widget `pdgw` (prepare_parameters(0, 0)) is iSWAP, 5.0 from SWAP --
    # through the device: widget -> solve_end -> MakhlinFunctionalCost
    from slam_decomposition_b200.utils.pd_playground import ParallelDrivenGateWidget
    undriven = ParallelDrivenGateWidget(N=10, gc=np.pi / 2, gg=0).solve_end()
    assert MakhlinFunctionalCost().unitary_fidelity(O.SWAP, undriven) == pytest.approx(5.0, abs=1e-7)
    assert O.J_T_LI(O.SWAP, O.ISWAP) == 5.0
    out = search_pulse(segs, O.SWAP, MakhlinFunctionalCost(), n_starts=4096, seed=3)
    assert out["loss"].shape == (4096,) and out["x"].shape == (4096, 6)
    # every reported loss is the oracle's cost of the reported parameters (the functional is quantised at 1e-8 per invariant)
    for i in list(range(0, 4096, 512)) + [int(np.argmin(out["loss"]))]:
        assert abs(out["loss"][i] - oracle_cost(out["x"][i])) < 1e-6
    p0 = np.array([np.pi / 2, 0, 0, 0, 2 * np.pi, 0])  # the notebook's starting point
    from_p0 = search_pulse(segs, O.SWAP, MakhlinFunctionalCost(), x0=p0[None])
    assert from_p0["best_loss"] <= oracle_cost(p0) + 1e-9
    assert out["best_loss"] <= from_p0["best_loss"] + 1e-9 and out["best_loss"] < 0.5  # thousands of starts beat the single one
