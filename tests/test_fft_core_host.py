"""The kernels' FFT arithmetic (csrc/fft_core.h, host/device code) compiled with g++ and run
thread-by-thread on the CPU: checks the 16 x 16 x 4 index algebra, twiddles, Hermitian
separation and the inverse packing against numpy's FFT."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "audio-style-transfer_b200", "csrc")


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    out = tmp_path_factory.mktemp("emul") / "libhost_emul.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                           os.path.join(CSRC, "host_emul.cpp"), "-o", str(out)])
    return ctypes.CDLL(str(out))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_complex_fft1024(emul):
    rng = np.random.default_rng(0)
    z = (rng.standard_normal(1024) + 1j * rng.standard_normal(1024)).astype(np.complex64)
    out = np.zeros(1024, dtype=np.complex64)
    bad = emul.emul_fft1024(_p(z), _p(out))
    assert bad == 0  # every output index written exactly once
    ref = np.fft.fft(z.astype(np.complex128))
    assert np.abs(out - ref).max() < 2e-6 * np.abs(ref).max()
    out2 = np.zeros(1024, dtype=np.complex64)
    assert emul.emul_fft1024_columns(_p(z), _p(out2)) == 0
    assert np.array_equal(out, out2)


def test_real_pair_forward(emul):
    rng = np.random.default_rng(1)
    fa = rng.standard_normal(1024).astype(np.float32)
    fb = rng.standard_normal(1024).astype(np.float32)
    xa = np.zeros(513, dtype=np.complex64)
    xb = np.zeros(513, dtype=np.complex64)
    emul.emul_rfft_pair(_p(fa), _p(fb), _p(xa), _p(xb))
    assert not np.isnan(xa.view(np.float32)).any() and not np.isnan(xb.view(np.float32)).any()
    ra, rb = np.fft.rfft(fa.astype(np.float64)), np.fft.rfft(fb.astype(np.float64))
    scale = max(np.abs(ra).max(), np.abs(rb).max())
    assert np.abs(xa - ra).max() < 2e-6 * scale and np.abs(xb - rb).max() < 2e-6 * scale
    # exact zeros where normalize() divides by std = 0 (dataloader.py:13)
    assert xa.imag[0] == 0 and xa.imag[512] == 0 and xb.imag[0] == 0 and xb.imag[512] == 0


def test_real_pair_inverse(emul):
    rng = np.random.default_rng(2)
    xa = (rng.standard_normal(513) + 1j * rng.standard_normal(513)).astype(np.complex64)
    xb = (rng.standard_normal(513) + 1j * rng.standard_normal(513)).astype(np.complex64)
    assert emul.emul_pack_check(_p(xa), _p(xb)) == 0
    fa = np.zeros(1024, dtype=np.float32)
    fb = np.zeros(1024, dtype=np.float32)
    emul.emul_irfft_pair(_p(xa), _p(xb), _p(fa), _p(fb))
    ra = np.fft.irfft(xa.astype(np.complex128), n=1024)  # ignores imag of DC / Nyquist like torch.istft
    rb = np.fft.irfft(xb.astype(np.complex128), n=1024)
    scale = max(np.abs(ra).max(), np.abs(rb).max())
    assert np.abs(fa - ra).max() < 2e-6 * scale and np.abs(fb - rb).max() < 2e-6 * scale


def test_warp_formulation_32x32(emul):
    """fft32 and the one-warp 1024-point transform (32 points per lane, one exchange tile) used by stft.cu / istft.cu."""
    rng = np.random.default_rng(3)
    z = (rng.standard_normal(32) + 1j * rng.standard_normal(32)).astype(np.complex64)
    out = np.zeros(32, dtype=np.complex64)
    emul.emul_fft32(_p(z), _p(out))
    ref = np.fft.fft(z.astype(np.complex128))
    assert np.abs(out - ref).max() < 1e-6 * np.abs(ref).max()
    z = (rng.standard_normal(1024) + 1j * rng.standard_normal(1024)).astype(np.complex64)
    out = np.zeros(1024, dtype=np.complex64)
    emul.emul_fft1024_warp(_p(z), _p(out))
    ref = np.fft.fft(z.astype(np.complex128))
    assert np.abs(out - ref).max() < 2e-6 * np.abs(ref).max()
    # two real frames in one transform: bins k and N - k separate exactly, imag DC / Nyquist exactly 0
    fa, fb = rng.standard_normal(1024).astype(np.float32), rng.standard_normal(1024).astype(np.float32)
    zz = (fa + 1j * fb).astype(np.complex64)
    emul.emul_fft1024_warp(_p(zz), _p(out))
    k = np.arange(513)
    zk, zp = out[k], out[(1024 - k) % 1024]
    xa = 0.5 * ((zk.real + zp.real) + 1j * (zk.imag - zp.imag))
    xb = 0.5 * ((zk.imag + zp.imag) + 1j * (zp.real - zk.real))
    ra, rb = np.fft.rfft(fa.astype(np.float64)), np.fft.rfft(fb.astype(np.float64))
    scale = max(np.abs(ra).max(), np.abs(rb).max())
    assert np.abs(xa - ra).max() < 2e-6 * scale and np.abs(xb - rb).max() < 2e-6 * scale
    assert xa.imag[0] == 0 and xa.imag[512] == 0 and xb.imag[0] == 0 and xb.imag[512] == 0
