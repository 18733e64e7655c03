"""CPU restatement of ``librosa.cqt`` as the reference calls it.  TEST INFRASTRUCTURE ONLY.

Reference call site: ``utilityFunctions.py:52``::

    cqt = librosa.cqt(waveform, sr=sample_rate, n_bins=n_bins, hop_length=hop_length)

with ``sr=22050, n_bins=84, hop_length=256`` (``utilityFunctions.py:39``) and every
other argument at its librosa default.  ``librosa`` is a third-party dependency that
is NOT vendored under ``/root/reference`` and not installed in this image (nor is
``soxr``, its default resampler), and the reference pins no version
(``README.md:160-168``).  PARITY UNPINNED: this file restates the documented
librosa >= 0.10 / 0.11 algorithm (``librosa.core.constantq.cqt -> vqt ->
__vqt_filter_fft / __cqt_response / __trim_stack``, ``librosa.filters.wavelet``,
``wavelet_lengths``, ``_relative_bandwidth``, ``util.sparsify_rows``,
``audio.resample(..., res_type="soxr_hq", scale=True)``); the only thing the
reference's own tests pin is the output shape ``(2, 862, 84)`` for a 220 500-sample
clip (``test_correctness.ipynb`` cell 3), which ``tests/test_oracle_cqt.py`` checks
together with known-answer tests (pure tones at bin centres, zero input, linearity).

The one part that cannot be restated bit-for-bit is libsoxr's "HQ" 2:1 decimator.
Its *specification* is restated from soxr's quality recipe (20-bit precision,
pass-band end ``1 - 0.05 / TO_3dB(rej)`` = 0.9136 x new Nyquist, stop-band begin
1.0 x new Nyquist, linear phase, ``(bits + 1) * 6.02`` dB = 126.4 dB rejection,
Kaiser-windowed sinc as in ``lsx_design_lpf`` / ``lsx_make_lpf``) and frozen here as
``decimator_taps()``; the CUDA path uses the same taps (``ast_host_decimator_taps``
in the C-ABI re-derives them; ``tests/test_cabi_host.py::test_plan_constants_match_oracle`` compares the two).
"""
from __future__ import annotations

import functools
import math

import numpy as np

# ---------------------------------------------------------------------------------
# geometry of the reference call (utilityFunctions.py:39, :52)
# ---------------------------------------------------------------------------------
SR = 22050
N_BINS = 84
BINS_PER_OCTAVE = 12
HOP = 256
FMIN = 32.70319566257483  # librosa.note_to_hz("C1")
N_OCTAVES = 7
WINDOW_BANDWIDTH_HANN = 1.50018310546875  # librosa.filters.WINDOW_BANDWIDTHS["hann"]


# ---------------------------------------------------------------------------------
# librosa.filters._relative_bandwidth / wavelet_lengths / wavelet
# ---------------------------------------------------------------------------------
def cqt_frequencies(n_bins: int = N_BINS, fmin: float = FMIN, bins_per_octave: int = BINS_PER_OCTAVE):
    """``librosa.interval_frequencies(intervals="equal")``: ``fmin * 2**(k / bpo)``."""
    return fmin * 2.0 ** (np.arange(n_bins, dtype=np.float64) / bins_per_octave)


def relative_bandwidth(freqs: np.ndarray) -> np.ndarray:
    """``librosa.filters._relative_bandwidth`` (centred log-frequency differences)."""
    logf = np.log2(freqs)
    bpo = np.empty_like(freqs)
    bpo[0] = 1.0 / (logf[1] - logf[0])
    bpo[-1] = 1.0 / (logf[-1] - logf[-2])
    bpo[1:-1] = 2.0 / (logf[2:] - logf[:-2])
    return (2.0 ** (2.0 / bpo) - 1.0) / (2.0 ** (2.0 / bpo) + 1.0)


def wavelet_lengths(freqs: np.ndarray, sr: float, alpha: np.ndarray, filter_scale: float = 1.0, gamma: float = 0.0):
    """``librosa.filters.wavelet_lengths``: fractional filter lengths and the cut-off frequency."""
    Q = float(filter_scale) / alpha
    filter_cutoff = np.max(freqs * (1 + 0.5 * WINDOW_BANDWIDTH_HANN / Q) + 0.5 * gamma)
    lengths = Q * sr / (freqs + gamma / alpha)
    return lengths, filter_cutoff


def _hann_periodic(n: int) -> np.ndarray:
    """``scipy.signal.get_window("hann", n, fftbins=True)``."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def wavelet_basis(freqs_oct: np.ndarray, sr: float, alpha_oct: np.ndarray):
    """``librosa.filters.wavelet(..., norm=1, pad_fft=True, window="hann")`` for one octave.

    Returns ``(basis (n_filters, n_fft) complex128, lengths)``; time support of filter
    ``k`` is ``arange(-ilen // 2, ilen // 2)`` (float floor division), centre-padded.
    """
    lengths, _ = wavelet_lengths(freqs_oct, sr, alpha_oct)
    filters = []
    for ilen, freq in zip(lengths, freqs_oct):
        n = np.arange(-ilen // 2, ilen // 2, dtype=np.float64)
        sig = np.exp(1j * (n * 2.0 * np.pi * freq / sr))  # util.phasor
        sig = sig * _hann_periodic(len(sig))  # __float_window("hann")(len(sig))
        sig = sig / np.sum(np.abs(sig))  # util.normalize(norm=1)
        filters.append(sig)
    max_len = int(2.0 ** (np.ceil(np.log2(max(lengths)))))
    basis = np.zeros((len(filters), max_len), dtype=np.complex128)
    for i, filt in enumerate(filters):
        lpad = (max_len - len(filt)) // 2  # util.pad_center
        basis[i, lpad : lpad + len(filt)] = filt
    return basis, lengths


def sparsify_rows(x: np.ndarray, quantile: float = 0.01) -> np.ndarray:
    """``librosa.util.sparsify_rows`` returned dense: zero the smallest-magnitude entries
    of each row whose cumulative share of the row's L1 magnitude is below ``quantile``."""
    mags = np.abs(x)
    norms = np.sum(mags, axis=1, keepdims=True)
    mag_sort = np.sort(mags, axis=1)
    cumulative_mag = np.cumsum(mag_sort / norms, axis=1)
    threshold_idx = np.argmin(cumulative_mag < quantile, axis=1)
    out = np.zeros_like(x)
    for i, j in enumerate(threshold_idx):
        keep = mags[i] >= mag_sort[i, j]
        out[i, keep] = x[i, keep]
    return out


def octave_fft_basis(octave: int, sr: float = SR, sparsity: float = 0.01):
    """``__vqt_filter_fft`` for octave ``octave`` (0 = top) followed by the
    ``fft_basis *= sqrt(sr / my_sr)`` rescale of ``vqt``.  Returns ``(fft_basis (12, 129), n_fft)``."""
    freqs = cqt_frequencies()
    alpha = relative_bandwidth(freqs)
    n_filters = BINS_PER_OCTAVE
    if octave == 0:
        sl = slice(-n_filters, None)
    else:
        sl = slice(-n_filters * (octave + 1), -n_filters * octave)
    my_sr = sr / 2.0**octave
    basis, lengths = wavelet_basis(freqs[sl], my_sr, alpha[sl])
    n_fft = basis.shape[1]
    basis = basis * (lengths[:, None] / float(n_fft))
    fft_basis = np.fft.fft(basis, n=n_fft, axis=1)[:, : n_fft // 2 + 1]
    fft_basis = sparsify_rows(fft_basis, quantile=sparsity)
    fft_basis = fft_basis * np.sqrt(sr / my_sr)
    return fft_basis, n_fft


# ---------------------------------------------------------------------------------
# soxr-HQ-like 2:1 decimator (see module docstring)
# ---------------------------------------------------------------------------------
def _bessel_i0(x):
    x = np.asarray(x, dtype=np.float64)
    term = np.ones_like(x)
    total = np.ones_like(x)
    y = x * x / 4.0
    for k in range(1, 64):
        term = term * y / (k * k)
        total = total + term
    return total


@functools.lru_cache(maxsize=None)
def decimator_spec():
    """Design parameters of the frozen decimator, restated from soxr's HQ recipe."""
    bits = 20.0
    db2 = 20.0 * math.log10(2.0)
    rej = bits * db2
    to_3db = (1.6e-6 * rej - 7.5e-4) * rej + 0.646
    passband_end = 1.0 - 0.05 / to_3db  # x new Nyquist
    stopband_begin = 1.0
    att = (bits + 1.0) * db2
    # normalise so that the INPUT Nyquist is 1 (decimation by 2 halves everything)
    fp, fs = passband_end / 2.0, stopband_begin / 2.0
    tr_bw = 0.5 * (fs - fp)
    fc = fs - tr_bw
    beta = 0.1102 * (att - 8.7)
    n = int(math.ceil((att - 7.95) / (2.285 * math.pi * (fs - fp)) + 1.0))
    if n % 2 == 0:
        n += 1  # odd length -> integer group delay -> zero-phase decimation
    return dict(passband_end=passband_end, att_db=att, fc=fc, beta=beta, num_taps=n, rho=0.5)


@functools.lru_cache(maxsize=None)
def _decimator_taps_cached():
    spec = decimator_spec()
    n, fc, beta, rho = spec["num_taps"], spec["fc"], spec["beta"], spec["rho"]
    m = n - 1
    z = np.arange(n, dtype=np.float64) - 0.5 * m
    with np.errstate(invalid="ignore", divide="ignore"):
        h = np.where(z == 0.0, fc, np.sin(fc * np.pi * z) / (np.pi * z))
    y = z / (0.5 * m + rho)
    h = h * _bessel_i0(beta * np.sqrt(1.0 - y * y)) / _bessel_i0(beta)
    h = h / np.sum(h)  # unit DC gain
    return h


def decimator_taps() -> np.ndarray:
    """Odd-length linear-phase low-pass FIR (float64, unit DC gain) used for every 2:1 stage."""
    return _decimator_taps_cached().copy()


def decimate2_direct(y: np.ndarray) -> np.ndarray:
    """Literal form of ``decimate2`` (``np.convolve``), kept to cross-check the fast one."""
    h = _decimator_taps_cached()
    half = (len(h) - 1) // 2
    n_out = (len(y) + 1) // 2
    ypad = np.concatenate([np.zeros(half), np.asarray(y, dtype=np.float64), np.zeros(half + 1)])
    # out[j] = sum_i h[i] * y[2j + i - half]  == correlate(ypad, h)[2j]  (h symmetric)
    full = np.convolve(ypad, h, mode="valid")
    return full[: 2 * n_out : 2] * math.sqrt(2.0)


def decimate2(y: np.ndarray) -> np.ndarray:
    """``librosa.resample(y, orig_sr=2, target_sr=1, res_type="soxr_hq", scale=True)``:
    output ``j`` is the low-passed input at time ``2j`` (zero-phase, zero-extension at
    both ends), length ``ceil(len / 2)``, then divided by ``sqrt(0.5)``.
    Polyphase evaluation (only the kept samples are computed), same arithmetic as the direct form."""
    from scipy.signal import upfirdn

    h = _decimator_taps_cached()
    half = (len(h) - 1) // 2  # even (192): conv(h, y)[2j + half] = upfirdn(...)[j + half // 2]
    n_out = (len(y) + 1) // 2
    out = upfirdn(h, np.asarray(y, dtype=np.float64), up=1, down=2)
    out = out[half // 2 : half // 2 + n_out]
    if len(out) < n_out:
        out = np.concatenate([out, np.zeros(n_out - len(out))])
    return out * math.sqrt(2.0)


# ---------------------------------------------------------------------------------
# librosa.core.constantq.vqt / __cqt_response / __trim_stack
# ---------------------------------------------------------------------------------
def _stft_ones_window(y: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """``librosa.stft(y, n_fft, hop_length=hop, window="ones", center=True,
    pad_mode="constant")`` -> ``(1 + n_fft // 2, 1 + len(y) // hop)``."""
    ypad = np.concatenate([np.zeros(n_fft // 2), y, np.zeros(n_fft // 2)])
    n_frames = 1 + len(y) // hop
    idx = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = ypad[idx]  # (frames, n_fft)
    return np.fft.rfft(frames, axis=1).T


def octave_signals(y: np.ndarray):
    """The seven signals the octave loop sees: ``y`` decimated 0..6 times."""
    sigs = [np.asarray(y, dtype=np.float64)]
    for _ in range(N_OCTAVES - 1):
        sigs.append(decimate2(sigs[-1]))
    return sigs


def cqt(y: np.ndarray) -> np.ndarray:
    """``librosa.cqt(y, sr=22050, n_bins=84, hop_length=256)`` -> ``(84, 1 + len(y) // 256)`` complex128."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    freqs = cqt_frequencies()
    alpha = relative_bandwidth(freqs)
    lengths, filter_cutoff = wavelet_lengths(freqs, SR, alpha)
    assert filter_cutoff < SR / 2.0
    resp = []
    my_y, my_hop = y, HOP
    for i in range(N_OCTAVES):
        fft_basis, n_fft = octave_fft_basis(i)
        D = _stft_ones_window(my_y, n_fft, my_hop)
        resp.append(fft_basis @ D)
        if my_hop % 2 == 0:
            my_hop //= 2
            my_y = decimate2(my_y)
    max_col = min(r.shape[-1] for r in resp)
    V = np.empty((N_BINS, max_col), dtype=np.complex128)
    end = N_BINS
    for r in resp:  # __trim_stack
        n_oct = r.shape[0]
        V[end - n_oct : end, :] = r[:, :max_col]
        end -= n_oct
    V /= np.sqrt(lengths)[:, None]  # scale=True
    return V


def get_CQT(waveform, sample_rate=22050, n_bins=84, hop_length=256) -> np.ndarray:
    """Restatement of ``utilityFunctions.get_CQT`` (``utilityFunctions.py:39-60``):
    ``(2, T, 84)`` float32, channel 0 real / channel 1 imaginary."""
    assert (sample_rate, n_bins, hop_length) == (SR, N_BINS, HOP), "oracle is frozen to the reference geometry"
    y = np.asarray(waveform).squeeze()
    V = cqt(y)
    out = np.stack([V.real, V.imag], axis=-1)  # (freq, time, 2)
    return np.transpose(out, (2, 1, 0)).astype(np.float32)


# ---------------------------------------------------------------------------------
# equivalent time-domain projection (what the CUDA kernels evaluate)
# ---------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def _time_kernel_cached():
    fft_basis, n_fft = octave_fft_basis(0)
    # resp[b, t] = sum_f basis[b, f] * sum_n frame[n] e^{-2 pi i f n / n_fft}
    n = np.arange(n_fft)
    f = np.arange(n_fft // 2 + 1)
    dft = np.exp(-2j * np.pi * np.outer(f, n) / n_fft)  # (129, 256)
    return fft_basis @ dft  # (12, 256) complex


def time_domain_kernel() -> np.ndarray:
    """``(12, 256)`` complex128 ``K`` with ``fft_basis @ rfft(frame) == K @ frame`` for the top
    octave; octave ``i`` uses ``K * sqrt(2**i)`` on the ``i``-times-decimated signal (the wavelet
    samples are identical in every octave because ``freq / my_sr`` is)."""
    return _time_kernel_cached().copy()
