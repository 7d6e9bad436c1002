"""Synthetic spectra of SURVEY.md §8d (configs 3 and 5): blood_01 axis geometry, K Lorentzians
plus N(0, sigma) noise, NumPy Generator(PCG64(20260000 + s)) per spectrum, draw order
maxp[K], hw[K], A[K], noise[N]."""
import numpy as np

X_MAX = 14.81146
X_WIDTH = 20.0236139622347
SIGNAL_BOUNDARIES = (11.8, -2.2)  # ordered for the decreasing axis


def axis(n: int) -> np.ndarray:
    i = np.arange(n, dtype=np.float64)
    return X_MAX - i * X_WIDTH / (float(n) - 1.0)


def spectrum(s: int, n: int = 131072, k: int = 500, hw_range=(5e-4, 3e-3), a_range=(1e4, 1e7),
             sigma: float = 300.0, integer: bool = False, x: np.ndarray | None = None) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(20260000 + s))
    if x is None:
        x = axis(n)
    maxp = rng.uniform(-2.0, 11.6, k)
    hw = np.exp(rng.uniform(np.log(hw_range[0]), np.log(hw_range[1]), k))
    amp = np.exp(rng.uniform(np.log(a_range[0]), np.log(a_range[1]), k))
    noise = rng.normal(0.0, sigma, n)
    y = np.zeros(n, dtype=np.float64)
    hw2 = hw * hw
    step = 64
    for j0 in range(0, k, step):
        d = x[:, None] - maxp[None, j0:j0 + step]
        y += (amp[None, j0:j0 + step] * hw2[None, j0:j0 + step] / (hw2[None, j0:j0 + step] + d * d)).sum(axis=1)
    y += noise
    if integer:
        y = np.rint(y)
    return y


def config3(s: int, n: int = 131072, integer: bool = False, x=None) -> np.ndarray:
    return spectrum(s, n=n, k=500, hw_range=(5e-4, 3e-3), integer=integer, x=x)


def config5(s: int, n: int = 131072, integer: bool = False, x=None) -> np.ndarray:
    return spectrum(s, n=n, k=3000, hw_range=(3e-4, 1.5e-3), integer=integer, x=x)
