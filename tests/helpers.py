import numpy as np


def image_metrics(a, b):
    """RMSE over RGB and the fraction of channels within |a-b| <= 1e-3 + 1e-3*|b|."""
    a = np.asarray(a, np.float64)[..., :3]
    b = np.asarray(b, np.float64)[..., :3]
    diff = np.abs(a - b)
    rmse = float(np.sqrt(np.mean((a - b) ** 2)))
    within = float(np.mean(diff <= 1e-3 + 1e-3 * np.abs(b)))
    return rmse, within


def random_rays(n, seed, scale=1.0, tmax=3.0e38):
    rng = np.random.default_rng(seed)
    ro = (rng.uniform(-1, 1, (n, 3)) * scale).astype(np.float32)
    rd = rng.normal(size=(n, 3)).astype(np.float32)
    rd /= np.linalg.norm(rd, axis=1, keepdims=True)
    return np.concatenate([ro, np.zeros((n, 1), np.float32), rd, np.full((n, 1), tmax, np.float32)], 1).astype(np.float32)
