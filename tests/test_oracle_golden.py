"""CPU: pins the oracle (oracle/exenv_oracle.c) on every golden vector the reference ships for this path
(reference tests/envs/<env>/data/, replay rule of reference tests/envs/<env>/test_<env>.py::test_step_results), plus
the known-answer tests of SURVEY.md Appendix D that need neither JAX nor a GPU."""
import numpy as np
import pytest

from helpers import ENVS, golden_oracle
from oracle import oracle as O


def replay(orc, g, n=None):
    st = orc.reset(1, mode=2, values=g["observations"][0:1])
    assert np.array_equal(st["obs"], g["observations"][0:1])  # D8: obs -> state -> obs round trip is exact
    state, gen = st["state"], [g["observations"][0]]
    n = n or g["actions"].shape[0]
    for i in range(n):
        r = orc.step(state, g["actions"][i:i + 1])
        state = r["state"]
        gen.append(r["obs"][0])
    return np.array(gen), g["observations"][: n + 1]


# agreement reached by an independent fp64 restatement
GOLDEN_ABS = {"pendulum": 3e-16, "mass_spring_damper": 3e-16, "fluid_tank": 0.0, "cart_pole": 3e-16, "acrobot": 1e-12}


@pytest.mark.parametrize("env", ENVS)
def test_oracle_reproduces_reference_golden_vectors(env):
    orc, g = golden_oracle(env)
    gen, ref = replay(orc, g)
    rtol = 1e-8 if env == "pmsm" else 1e-16
    assert np.allclose(gen, ref, rtol=rtol, atol=1e-8)  # the reference's own assertion
    if env in GOLDEN_ABS:
        assert np.max(np.abs(gen - ref)) <= GOLDEN_ABS[env]
    else:  # PMSM: unstable closed loop, observations reach 6e7
        assert np.max(np.abs(gen - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-10


def test_pmsm_hexagon_constants_are_float32_rounded():
    """ROTATION_MAP is complex64; exact fp64 sqrt(3)/2 must break the golden replay."""
    import subprocess, tempfile, ctypes as C
    from pathlib import Path

    src = (Path(O.__file__).parent / "exenv_oracle_impl.h").read_text()
    assert "0.8660254037844386f" in src
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        for name in ("exenv_oracle.c", "exenv_oracle.h"):
            (td / name).write_text((Path(O.__file__).parent / name).read_text())
        (td / "exenv_oracle_impl.h").write_text(src.replace("(REAL)0.8660254037844386f", "(REAL)0.8660254037844386")
                                                .replace("(REAL)0.5f", "(REAL)0.5"))
        so = td / "bad.so"
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", str(so),
                        str(td / "exenv_oracle.c"), "-lm"], check=True)
        bad = C.CDLL(str(so))
        orc, g = golden_oracle("pmsm")
        good_lib = O._lib
        try:
            O._lib = bad
            bad.exo_sim_ahead_f64.restype = C.c_long
            gen, ref = replay(orc, g)
        finally:
            O._lib = good_lib
        assert not np.allclose(gen, ref, rtol=1e-8, atol=1e-8)


def test_tsit5_tableau_and_dense_output_consistency():
    """row sums = c, order conditions to order 5, b_i(1) = b_i, sum b_i(theta) = theta."""
    c = np.array([0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0])
    a = [[], [0.161], [-0.008480655492356989, 0.335480655492357],
         [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
         [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
         [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
         [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]]
    for i in range(1, 7):
        assert abs(sum(a[i]) - c[i]) < 1e-15
    b = np.array(a[6] + [0.0])
    A = np.zeros((7, 7))
    for i in range(7):
        A[i, :len(a[i])] = a[i]
    assert abs(b.sum() - 1) < 1e-15
    for k in range(1, 5):
        assert abs(b @ c ** k - 1 / (k + 1)) < 1e-15
    assert abs(b @ (A @ c) - 1 / 6) < 1e-15 and abs(b @ (A @ c ** 2) - 1 / 12) < 1e-15
    assert abs(b @ (A @ (A @ c)) - 1 / 24) < 1e-15 and abs((b * c) @ (A @ c) - 1 / 8) < 1e-15

    def bw(t):
        return np.array([
            -1.0530884977290216 * t * (t - 1.3299890189751412) * (t ** 2 - 1.4364028541716351 * t + 0.7139816917074209),
            0.1017 * t ** 2 * (t ** 2 - 2.1966568338249754 * t + 1.2949852507374631),
            2.490627285651252793 * t ** 2 * (t ** 2 - 2.38535645472061657 * t + 1.57803468208092486),
            -16.54810288924490272 * (t - 1.21712927295533244) * (t - 0.61620406037800089) * t ** 2,
            47.37952196281928122 * (t - 1.203071208372362603) * (t - 0.658047292653547382) * t ** 2,
            -34.87065786149660974 * (t - 1.2) * (t - 0.666666666666666667) * t ** 2,
            2.5 * (t - 1.0) * (t - 0.6) * t ** 2])

    assert np.max(np.abs(bw(1.0) - b)) < 3e-15
    for t in (0.0, 0.13, 0.5, 0.77, 1.0):
        assert abs(bw(t).sum() - t) < 3e-15


def test_msd_solvers_against_closed_forms():
    """linear mass-spring-damper, constant force: Tsit5 vs expm, Euler vs (I + hA)^N."""
    from scipy.linalg import expm

    k, d, m, F, h, T = 100.0, 1.0, 1.0, 7.3, 1e-4, 2000
    A = np.array([[0, 1], [-k / m, -d / m]])
    yss, y0 = np.array([F / k, 0.0]), np.array([0.5, -1.0])
    acts = np.full((1, T, 1), F / 20.0)
    st = [np.array([0.5]), np.array([-1.0])]
    r = O.Oracle("mass_spring_damper", solver=O.TSIT5).sim_ahead(st, acts, h, h)
    ana = np.array([yss + expm(A * t) @ (y0 - yss) for t in np.arange(T + 1) * h])
    num = np.stack([r["traj"][0][0], r["traj"][1][0]], 1)
    assert np.max(np.abs(num - ana)) < 1e-13
    e = O.Oracle("mass_spring_damper", solver=O.EULER)
    s = st
    for i in range(T):
        s = e.step(s, acts[:, i])["state"]
    M = np.linalg.matrix_power(np.eye(2) + h * A, T)
    ref = yss + M @ (y0 - yss)
    assert abs(s[0][0] - ref[0]) < 1e-12 and abs(s[1][0] - ref[1]) < 1e-12


@pytest.mark.parametrize("motor", ["BRUSA", "SEW"])
def test_bilinear_lut_against_scipy(motor):
    """the oracle's LUT path == scipy RegularGridInterpolator(bounds_error=False, fill_value=None)."""
    from scipy.interpolate import RegularGridInterpolator

    gx, gy, maps = O.load_motor_lut(motor)
    rng = np.random.default_rng(0)
    B = 500
    i_d = rng.uniform(gx[0] - 2 * (gx[1] - gx[0]), gx[-1] + 2 * (gx[1] - gx[0]), B)
    i_q = rng.uniform(gy[0] - 2 * (gy[1] - gy[0]), gy[-1] + 2 * (gy[1] - gy[0]), B)
    psi_d = RegularGridInterpolator((gx, gy), maps[4], bounds_error=False, fill_value=None)(np.stack([i_d, i_q], 1))
    psi_q = RegularGridInterpolator((gx, gy), maps[5], bounds_error=False, fill_value=None)(np.stack([i_d, i_q], 1))
    orc = O.Oracle("pmsm", saturated=True, motor=motor)
    p = orc.static["p"]
    torque_ref = 1.5 * p * (psi_d * i_q - psi_q * i_d)
    # torque of the state after a zero-length... use the reset-from-state path: observe torque via one Euler step is
    # state dependent, so evaluate the oracle's torque through a step with tau -> 0
    tiny = O.Oracle("pmsm", saturated=True, motor=motor, tau=1e-300)
    z = np.zeros(B)
    st = [z, z, z, i_d, i_q, z, z]
    out = tiny.step(st, np.zeros((B, 2)))["state"]
    assert np.max(np.abs(out[5] - torque_ref)) <= 1e-9 * (1 + np.max(np.abs(torque_ref)))


def test_sim_ahead_edge_cases_and_errors():
    orc = O.Oracle("pendulum", solver=O.TSIT5)
    st = orc.reset(3)["state"]
    one = orc.sim_ahead(st, np.zeros((3, 1, 1)), 1e-4, 1e-4)
    assert one["obs"].shape == (3, 2, 2) and one["n_steps"] == 1
    with pytest.raises(RuntimeError, match="maximum number of solver steps"):
        orc.sim_ahead(st, np.zeros((3, 5000, 1)), 1e-4, 1e-4)
    big = orc.sim_ahead(st, np.zeros((3, 5000, 1)), 1e-4, 1e-4, max_steps=1 << 30)
    assert big["obs"].shape == (3, 5001, 2)
