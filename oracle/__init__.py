"""CPU oracle for the SLAM template-evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``slam_decomposition_b200``; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it, and
there only as the checker / timed CPU baseline, never as the thing shipped.

Parity status: the reference's own test-suite pins nothing
(``/root/reference/src/tests/main_test.py:4-6`` is ``assert 1 == 1``) and its
numerics live in qiskit / qutip / weylchamber, none of which are installed
here, so the reference cannot be executed in this image.  The oracle is a
numpy/scipy restatement of the reference call sites and of the published
weylchamber algorithm; it IS pinned against every recorded output the
reference tree holds for this path (notebook logs, README, analytic chamber
points: ``tests/golden/kats.json``, ``tests/test_oracle_kats.py``).
"""
from .slam_oracle import *  # noqa: F401,F403
