"""Pins the CPU oracle to the known-answer tables and the H2 fixture of SURVEY.md 8(c)."""
import numpy as np
import pytest

RHO = np.array([1e-6, 1e-3, 5e-2, 0.3, 1.0, 10.0, 150.0])
SIG = np.array([1e-14, 1e-6, 1e-2, 0.2, 5.0, 100.0, 2e4])

KAT = {
    0: dict(exc=[-1.216220516826753e-08, -9.872067156718419e-05, -1.602293898963634e-02, -1.668678991594554e-01,
                 -8.101513786888132e-01, -1.682816332703021e+01, -6.061593077558503e+02],
            vrho=[-1.744840741064480e-02, -1.299350417367609e-01, -4.195636808841949e-01, -7.300201075731788e-01,
                  -1.065633005632815e+00, -2.222938090747775e+00, -5.359170230213710e+00]),
    1: dict(exc=[-1.168269025884942e-08, -1.041991057278235e-04, -1.660469496747921e-02, -1.673946834147352e-01,
                 -8.141235302515013e-01, -1.682351326575120e+01, -6.060870452200159e+02],
            vrho=[-1.659840344714811e-02, -1.210364875264344e-01, -3.937973752969727e-01, -7.224504887825094e-01,
                  -1.051764936862050e+00, -2.221492813126824e+00, -5.358216445054708e+00],
            vsigma=[+1.915555801286228e+04, -7.846134570727507e+00, -8.368311313086603e-02, -5.335001032377830e-03,
                    -1.463594244973456e-03, -1.109163118280122e-05, -1.159699526254412e-07]),
    2: dict(exc=[-1.089664865127960e-08, -9.108499938854123e-05, -1.384582801590466e-02, -1.361680127939420e-01,
                 -6.620581907571683e-01, -1.341949242215582e+01, -4.825996024456696e+02],
            vrho=[-1.095068506660968e-02, -9.333543629284999e-02, -3.210190834429153e-01, -5.716949888005017e-01,
                  -8.327398616254684e-01, -1.766745916842922e+00, -4.265455784383697e+00],
            vsigma=[-1.258266224297871e+05, -9.201723571820336e+00, -8.075965010741540e-02, -1.347725969665545e-02,
                    -2.821598403250905e-03, -1.682710621266796e-04, -4.710494589865961e-06]),
}


@pytest.mark.parametrize("xc", [0, 1, 2])
def test_pointwise_kat_compat(oracle, xc):
    exc, vr, vs = oracle.functional_points(xc, RHO, SIG, mode=oracle.COMPAT)
    np.testing.assert_allclose(exc, KAT[xc]["exc"], rtol=1e-12)
    np.testing.assert_allclose(vr, KAT[xc]["vrho"], rtol=1e-12)
    if xc:
        np.testing.assert_allclose(vs, KAT[xc]["vsigma"], rtol=1e-12)


def test_lda_exact_derivative_reference_values(oracle):
    # exact-derivative v at rho = 1 and 1e-2
    _, v, _ = oracle.functional_points(0, [1.0, 1e-2], mode=oracle.EXACT)
    np.testing.assert_allclose(v, [-1.0646834050, -0.2560295400], rtol=2e-10)


def _fd(oracle, xc, mode, rho, sig):
    f = lambda r, s: oracle.functional_points(xc, r, s, mode=mode)[0]
    h = 1e-4 * rho
    dr = (-f(rho + 2 * h, sig) + 8 * f(rho + h, sig) - 8 * f(rho - h, sig) + f(rho - 2 * h, sig)) / (12 * h)
    hs = 1e-4 * sig
    ds = (-f(rho, sig + 2 * hs) + 8 * f(rho, sig + hs) - 8 * f(rho, sig - hs) + f(rho, sig - 2 * hs)) / (12 * hs)
    return dr, ds


@pytest.mark.parametrize("xc", [0, 1, 2])
def test_exact_mode_is_derivative_of_energy(oracle, xc):
    rho, sig = RHO[1:], SIG[1:]
    _, vr, vs = oracle.functional_points(xc, rho, sig, mode=oracle.EXACT)
    dr, ds = _fd(oracle, xc, oracle.EXACT, rho, sig)
    np.testing.assert_allclose(vr, dr, rtol=1e-8, atol=1e-10)
    if xc:
        np.testing.assert_allclose(vs, ds, rtol=1e-6)


def test_compat_mode_reproduces_documented_deviations(oracle):
    """D1 (VWN5) and D2 (PBE-c) make the reference potentials differ from the energy derivative."""
    rho, sig = RHO[1:], SIG[1:]
    _, v, _ = oracle.functional_points(0, rho, mode=oracle.COMPAT)
    dr, _ = _fd(oracle, 0, oracle.COMPAT, rho, sig)
    assert 4e-4 < np.max(np.abs(v - dr)) < 2e-3          # D1: 5e-4 .. 1.8e-3 Ha
    _, v, _ = oracle.functional_points(1, rho, sig, mode=oracle.COMPAT)
    dr, _ = _fd(oracle, 1, oracle.COMPAT, rho, sig)
    assert 1e-3 < np.max(np.abs(v - dr)) < 5e-3          # D2: up to 4.4e-3 Ha
    # B3LYP has no deviation
    e0 = oracle.functional_points(2, rho, sig, mode=oracle.COMPAT)
    e1 = oracle.functional_points(2, rho, sig, mode=oracle.EXACT)
    for a, b in zip(e0, e1):
        np.testing.assert_array_equal(a, b)


def test_row_gate(oracle):
    exc, vr, vs = oracle.functional_points(1, [5e-13, 2e-12], [1e-3, 1e-3], gate=True)
    assert exc[0] == 0 and vr[0] == 0 and vs[0] == 0
    assert exc[1] != 0


H2_KAT = {0: (-0.683240084985, -0.448744125728, -0.302576546237),
          1: (-0.714211888015, -0.463813769518, -0.311372142021),
          2: (-0.591838579199, -0.379855055638, -0.254425920412)}


@pytest.mark.parametrize("xc", [0, 1, 2])
def test_h2_fixture(oracle, h2_fixture, xc):
    mol, basis, coords, w, dm = h2_fixture
    ao, grad = oracle.eval_ao(coords, basis, deriv=1)
    S = np.einsum("g,gi,gj->ij", w, ao, ao)
    np.testing.assert_allclose(S, [[0.99999999, 0.67808713], [0.67808713, 0.99999999]], atol=1e-8)
    e, v, rho, _ = oracle.compute_xc(xc, dm, ao, w, grad, want_density=True)
    assert abs(np.dot(w, rho) - 2.0) < 2e-8
    sv = oracle.sym(v)
    assert abs(e - H2_KAT[xc][0]) < 1e-11
    assert abs(sv[0, 0] - H2_KAT[xc][1]) < 1e-11 and abs(sv[1, 1] - H2_KAT[xc][1]) < 1e-11
    assert abs(sv[0, 1] - H2_KAT[xc][2]) < 1e-11


def test_nonsymmetric_density_follows_reference_loops(oracle):
    """The reference's double loop uses D as given; rho depends only on sym(D), grad on D + D^T."""
    rng = np.random.default_rng(3)
    ao = rng.standard_normal((50, 5)); g = rng.standard_normal((3, 50, 5))
    D = rng.standard_normal((5, 5))
    r1, g1, _ = oracle.density(D, ao, g)
    r2, g2, _ = oracle.density(0.5 * (D + D.T), ao, g)
    np.testing.assert_allclose(r1, r2, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(g1, g2, rtol=1e-12, atol=1e-12)


def test_coulomb_restatement(oracle):
    rng = np.random.default_rng(5)
    n = 4
    eri = rng.standard_normal((n * n, n * n))
    dm = rng.standard_normal((n, n))
    J = oracle.coulomb(eri, dm)
    np.testing.assert_allclose(J.ravel(), eri.T @ dm.ravel(), rtol=1e-13)  # column-major Dgemv, no transpose
