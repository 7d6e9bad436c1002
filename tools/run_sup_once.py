"""ncu target: one superposition_vec launch (K8) of 2^22 points x 2 048 Lorentzians in the library's
current superposition mode (MDB_SUPERPOSITION=exact|fast)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from metabodecon_rust_b200.lorentzian import superposition_vec_array  # noqa: E402

rng = np.random.default_rng(1)
p, n = 2048, 1 << 22
hw = np.exp(rng.uniform(np.log(5e-4), np.log(3e-3), p))
sf = np.exp(rng.uniform(0.0, np.log(1e4), p))
lor = np.stack([sf * hw, hw * hw, rng.uniform(0.0, 10.0, p)], axis=1)
x = np.linspace(-2.2, 11.8, n)
out = superposition_vec_array(x, lor)
out = superposition_vec_array(x, lor)
print("checksum", float(out[::4096].sum()))
