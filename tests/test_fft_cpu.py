"""CPU validation of the FFT building blocks the CUDA kernels are made of (airwave_b200/csrc/aw_fft.cuh,
aw_fft_reg.cuh): the very same per-butterfly / per-pass functions are compiled with g++ and driven with loops
in place of threads (tests/cpu/*.cpp), then compared with numpy's float64 FFT.  This catches index, twiddle
and pass-plan mistakes without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

FP = C.POINTER(C.c_float)


def _build(name):
    src = os.path.join(ROOT, "tests", "cpu", name + ".cpp")
    out = os.path.join(ROOT, "tests", "cpu", "_build", name + ".so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src, os.path.join(ROOT, "airwave_b200", "csrc", "aw_fft.cuh"), os.path.join(ROOT, "airwave_b200", "csrc", "aw_fft_reg.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    return C.CDLL(out)


@pytest.fixture(scope="module")
def stockham():
    return _build("fft_harness")


@pytest.fixture(scope="module")
def regfft():
    return _build("fft_reg_harness")


@pytest.mark.parametrize("log2m", range(2, 14))
def test_register_radix_fft_matches_numpy(regfft, log2m):
    M = 1 << log2m
    rng = np.random.default_rng(log2m)
    z = rng.uniform(-1, 1, M) + 1j * rng.uniform(-1, 1, M)
    inp = np.stack([z.real, z.imag], 1).astype(np.float32).copy()
    out = np.zeros((M, 2), np.float32)
    assert regfft.harness_regfft(inp.ctypes.data_as(FP), log2m, out.ctypes.data_as(FP)) == 0
    ref = np.fft.fft(inp[:, 0].astype(np.float64) + 1j * inp[:, 1].astype(np.float64))
    assert np.abs((out[:, 0] + 1j * out[:, 1]) - ref).max() <= 4e-7 * np.abs(ref).max()


@pytest.mark.parametrize("log2m", range(2, 14))
def test_per_pass_twiddle_tables_give_bit_identical_transforms(regfft, log2m):
    """compute_pt (conflict-free per-pass tables, used by the persistent kernel) reads the very same twiddle values as
    compute (half-circle table): the transforms must agree bit for bit."""
    M = 1 << log2m
    rng = np.random.default_rng(200 + log2m)
    inp = rng.uniform(-1, 1, (M, 2)).astype(np.float32)
    a, b = np.zeros((M, 2), np.float32), np.zeros((M, 2), np.float32)
    regfft.harness_regfft_use_pt(0)
    assert regfft.harness_regfft(inp.ctypes.data_as(FP), log2m, a.ctypes.data_as(FP)) == 0
    regfft.harness_regfft_use_pt(1)
    try:
        assert regfft.harness_regfft(inp.ctypes.data_as(FP), log2m, b.ctypes.data_as(FP)) == 0
    finally:
        regfft.harness_regfft_use_pt(0)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("log2m", range(2, 14))
def test_twiddle_products_stay_within_the_transform_tolerance(regfft, log2m):
    """K2 / K4 read six twiddles per radix-16 pass and multiply for the other nine (compute_pt<P, true>): same tolerance against
    numpy as the table-only transform, and bit-identical to it where no pass is a radix-16 pass with twiddles."""
    M = 1 << log2m
    rng = np.random.default_rng(300 + log2m)
    inp = rng.uniform(-1, 1, (M, 2)).astype(np.float32)
    a, b = np.zeros((M, 2), np.float32), np.zeros((M, 2), np.float32)
    try:
        regfft.harness_regfft_use_pt(1)
        assert regfft.harness_regfft(inp.ctypes.data_as(FP), log2m, a.ctypes.data_as(FP)) == 0
        regfft.harness_regfft_use_pt(2)
        assert regfft.harness_regfft(inp.ctypes.data_as(FP), log2m, b.ctypes.data_as(FP)) == 0
    finally:
        regfft.harness_regfft_use_pt(0)
    ref = np.fft.fft(inp[:, 0].astype(np.float64) + 1j * inp[:, 1].astype(np.float64))
    assert np.abs((b[:, 0] + 1j * b[:, 1]) - ref).max() <= 4e-7 * np.abs(ref).max()
    if log2m in (2, 3, 4, 5, 6, 7, 9, 10):       # pass plans without a twiddled radix-16 pass (aw_fft_reg.cuh pass_log2r)
        assert np.array_equal(a, b)
    else:
        assert not np.array_equal(a, b)


@pytest.mark.parametrize("log2m", range(2, 14))
def test_real_fft_split_steps_match_numpy_and_round_trip(stockham, log2m):
    M, nf = 1 << log2m, 3
    N = 2 * M
    rng = np.random.default_rng(100 + log2m)
    x = rng.uniform(-1, 1, (nf, N)).astype(np.float32)
    spec = np.zeros((nf, M, 2), np.float32)
    ny = np.zeros(nf, np.float32)
    stockham.harness_rfft_forward(x.ctypes.data_as(FP), log2m, nf, spec.ctypes.data_as(FP), ny.ctypes.data_as(FP))
    X = 2 * np.fft.rfft(x.astype(np.float64), axis=1)     # vDSP convention: forward = 2 x DFT
    scale = np.abs(X).max()
    got = spec[..., 0] + 1j * spec[..., 1]
    assert np.abs(got[:, 1:] - X[:, 1:M]).max() <= 4e-7 * scale
    assert np.abs(got[:, 0].real - X[:, 0].real).max() <= 4e-7 * scale and np.all(got[:, 0].imag == 0)
    assert np.abs(ny - X[:, M].real).max() <= 4e-7 * scale
    back = np.zeros((nf, N), np.float32)
    stockham.harness_irfft(spec.ctypes.data_as(FP), ny.ctypes.data_as(FP), log2m, nf, back.ctypes.data_as(FP))
    assert np.abs(back / (2 * N) - x).max() <= 2e-6
