"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): every kernel once, tiny sizes."""
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import oracle as O
from helpers import make_pair, BASES
from slam_decomposition_b200 import engine
from slam_decomposition_b200.utils.gates import parallel_drive_volume as pdv
rng = np.random.default_rng(0)
V = torch.as_tensor(O.haar_unitary(rng, 8), device="cuda")
for kind, slots, kw in (("riswap", (0.5,), {}), ("cg", ("Q", 0.2, np.pi / 4, "Q", 0.5), {}), ("riswap", ("Q",), {"vz_only": True})):
    desc, orc = make_pair(kind, slots, k=3, **kw)
    X = torch.as_tensor(rng.uniform(0, 6, (70, orc.n_params)), device="cuda")
    for lpp in ("4", "2", "1"):
        os.environ["SLAM_B200_LPP"] = lpp
        engine.loss_grad(desc, X, V)
    engine.template_eval(desc, X)
    opts = engine.opt_defaults(); opts.max_iter = 40
    engine.lbfgs_solve(desc, V, 4, opts, seed=1)
desc, orc = make_pair("riswap", (0.5,), k=2)
nm = engine.nm_defaults(); nm.max_iter = 30; nm.cost_kind = 3
engine.nm_solve(desc, V, 2, nm, seed=2)
engine.weyl(V, want_g=True)
b = pdv.smush_template(*BASES["sqiSwap"], 2)
pdv.coverage_histogram(b, 2000, seed=3, nbins=16)
engine.pd_trajectory(torch.zeros((3, 8), device="cuda", dtype=torch.float64), torch.ones((3, 4), device="cuda", dtype=torch.float64),
                     torch.ones((3, 4), device="cuda", dtype=torch.float64), 0.1)
# round-1 additions: plain coverage (staged Philox), smush K1/K2 (adjoint), K5c in all gradient modes, chained sweep
pdv.coverage_histogram(pdv.plain_template(*BASES["CNOT"], 3), 3000, seed=4, nbins=16)
X = torch.as_tensor(rng.uniform(-2, 2, (70, b.desc.n_params)), device="cuda")
engine.template_eval(b.desc, X)
engine.loss_grad(b.desc, X, V)
engine.loss_grad(b.desc, X, V, want_grad=False)
for mode in (0, 1, 2):
    opts = engine.opt_defaults(); opts.max_iter = 6
    engine.fd_lbfgs_solve(b.desc, V, 3, opts, seed=5, central=mode)
from slam_decomposition_b200.basis import CircuitTemplate
from slam_decomposition_b200.cost_function import BasicCost
from slam_decomposition_b200.optimizer import TemplateOptimizer
from slam_decomposition_b200.utils.gates.custom_gates import RiSwapGate
opt = TemplateOptimizer(CircuitTemplate(base_gates=[RiSwapGate(0.5)], maximum_span_guess=3, preseed=False), BasicCost(),
                        override_fail=True, training_restarts=2)
o = engine.opt_defaults(); o.max_iter = 30
opt.approximate_targets(V.cpu().numpy(), range(1, 4), opts=o)
torch.cuda.synchronize()
print("sanitizer case done")
