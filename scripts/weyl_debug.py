import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
import oracle as O
from slam_decomposition_b200 import engine
rng = np.random.default_rng(1)
base = {"I": np.eye(4), "CNOT": O.CNOT, "SWAP": O.SWAP, "ISWAP": O.ISWAP, "sqiswap": O.riswap(0.5), "B": O.berkeley(),
        "sqCNOT": O.conversion_gain(0, 0, np.pi / 4, np.pi / 4, 0.5)}
for name, M in base.items():
    cm = O.fold_c1(O.c1c2c3_raw(M))
    Us = []
    for _ in range(50):
        k1 = np.kron(O.u3(*rng.uniform(0, 7, 3)), O.u3(*rng.uniform(0, 7, 3)))
        k2 = np.kron(O.u3(*rng.uniform(0, 7, 3)), O.u3(*rng.uniform(0, 7, 3)))
        Us.append(np.exp(1j * rng.uniform(0, 7)) * (k1 @ M @ k2))
    c, _ = engine.weyl(torch.as_tensor(np.stack(Us), device="cuda"), fold=True)
    c = c.cpu().numpy()
    err = np.abs(c - cm).max(axis=1)
    i = int(err.argmax())
    print(name, cm, "max err %.3e" % err.max(), "worst", c[i])
