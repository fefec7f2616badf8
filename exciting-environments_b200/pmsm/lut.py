"""Host-side, one-off preparation of the PMSM flux-map look-up tables (reference
pmsm/pmsm_env.py:316-363 ``generate_interpolators_and_lut``): NaN cells are filled with their nearest valid
neighbour (on index coordinates), the border is replicated by one cell so that linear extrapolation becomes
constant, and the padded regular grids are built.  The result is uploaded once per plan; the per-step bilinear
look-up itself runs inside the CUDA kernels (csrc/exenv_envs.cuh: lut_eval)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

QUANTITIES = ("L_dd", "L_dq", "L_qd", "L_qq", "Psi_d", "Psi_q")
DATA_DIR = Path(__file__).resolve().parent / "data"


def prepare(mat: dict):
    """Returns ``(gx, gy, maps)``: i_d grid (nx,), i_q grid (ny,), maps (6, nx, ny) float64 holding
    ``values[x_idx][y_idx]`` of the six quantities in ``QUANTITIES`` order."""
    from scipy.interpolate import griddata

    i_d_vec, i_q_vec = np.asarray(mat["i_d_vec"]), np.asarray(mat["i_q_vec"])
    d_lo, d_hi, q_lo, q_hi = i_d_vec.min(), i_d_vec.max(), i_q_vec.min(), i_q_vec.max()
    step_d = (d_hi - d_lo) / (i_d_vec.shape[1] - 1)
    step_q = (q_hi - q_lo) / (i_q_vec.shape[1] - 1)
    padded = []
    for name in QUANTITIES:
        grid = np.array(mat[name], copy=True)
        holes = np.isnan(grid)
        if holes.any():
            rows, cols = np.indices(grid.shape)
            grid[holes] = griddata((rows[~holes], cols[~holes]), grid[~holes], (rows[holes], cols[holes]),
                                   method="nearest")
        padded.append(np.pad(grid, 1, mode="edge"))
    n_q, n_d = padded[0].shape
    gx = np.linspace(d_lo - step_d, d_hi + step_d, n_d)
    gy = np.linspace(q_lo - step_q, q_hi + step_q, n_q)
    maps = np.stack([np.asarray(m.T, dtype=np.float64) for m in padded])
    return gx, gy, maps


def load(motor_name: str):
    from scipy.io import loadmat

    return prepare(loadmat(DATA_DIR / f"LUT_{motor_name}_jax_grad.mat"))


def bilinear(gx, gy, values, x, y):
    """NumPy bilinear interpolation with linear extrapolation on a regular grid (only used for the torque entry of
    randomly drawn reset states)."""
    x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
    i = np.clip(np.searchsorted(gx, x, side="left") - 1, 0, gx.size - 2)
    j = np.clip(np.searchsorted(gy, y, side="left") - 1, 0, gy.size - 2)
    wx = (x - gx[i]) / (gx[i + 1] - gx[i])
    wy = (y - gy[j]) / (gy[j + 1] - gy[j])
    return (values[i, j] * (1 - wx) * (1 - wy) + values[i, j + 1] * (1 - wx) * wy + values[i + 1, j] * wx * (1 - wy)
            + values[i + 1, j + 1] * wx * wy)
