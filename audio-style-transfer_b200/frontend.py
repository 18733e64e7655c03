"""Batched spectral front-end: a thin, typed Python face over the C-ABI (``include/ast_frontend.h``).

PyTorch is used for device memory and streams only; every transform below is one C-ABI call
into hand-written sm_100a kernels.  There is no CPU implementation here.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import LAYOUT_FLAT, LAYOUT_SECTIONS

N_FFT = 1024
HOP = 256
F_STFT = 513
F_CQT = 84
F_TOTAL = 597
WINDOW_SIZE = 287      # utilityFunctions.py:8
OVERLAP_FRAMES = 96    # utilityFunctions.py:10
SAMPLE_RATE = 22050

_LAYOUTS = {"flat": LAYOUT_FLAT, "sections": LAYOUT_SECTIONS}


def num_frames(n_samples: int) -> int:
    return int(_lib.load().ast_num_frames(int(n_samples)))


def num_sections(n_frames: int, window_size: int = WINDOW_SIZE, overlap_frames: int = OVERLAP_FRAMES) -> int:
    return int(_lib.load().ast_num_sections(int(n_frames), int(window_size), int(overlap_frames)))


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream_ptr(device: torch.device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class FrontEnd:
    """One plan (constant tables) on one CUDA device, plus a grow-only scratch buffer.

    ``window_size`` / ``overlap_frames`` are the section geometry (287 / 96 by default as in
    ``utilityFunctions.py:8-10``; the evaluation scripts use 287 / 86)."""

    def __init__(self, device=None, window_size: int = WINDOW_SIZE, overlap_frames: int = OVERLAP_FRAMES):
        if not torch.cuda.is_available():
            raise RuntimeError("audio-style-transfer_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"FrontEnd needs a CUDA device, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.window_size = int(window_size)
        self.overlap_frames = int(overlap_frames)
        cfg = _lib.AstConfig()
        _lib.check(self.lib.ast_default_config(ctypes.byref(cfg)))
        cfg.window_size = self.window_size
        cfg.overlap_frames = self.overlap_frames
        cfg.device = self.device.index
        plan = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ast_plan_create(ctypes.byref(cfg), ctypes.byref(plan)))
        self._plan = plan
        self._ws: Optional[torch.Tensor] = None
        self._resamplers: dict = {}

    def __del__(self):
        for r in getattr(self, "_resamplers", {}).values():
            try:
                self.lib.ast_resampler_destroy(r)
            except Exception:
                pass
        self._resamplers = {}
        plan = getattr(self, "_plan", None)
        if plan is not None and plan.value:
            try:
                self.lib.ast_plan_destroy(plan)
            except Exception:
                pass
            self._plan = None

    # ------------------------------------------------------------------ helpers
    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    def _wave(self, wave: torch.Tensor, lengths: Optional[torch.Tensor]) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        if wave.ndim == 1:
            wave = wave.unsqueeze(0)
        if wave.ndim != 2:
            raise ValueError(f"wave must be (B, L) or (L,), got {tuple(wave.shape)}")
        wave = wave.to(device=self.device, dtype=torch.float32)
        if wave.stride(-1) != 1 or (wave.shape[0] > 1 and wave.stride(0) < wave.shape[1]):
            wave = wave.contiguous()
        B, L = wave.shape
        if wave.data_ptr() % 16 or (B > 1 and wave.stride(0) % 2):
            # the decimator reads sample pairs: every row must start on an even (8-byte aligned) offset
            padded = torch.zeros((B, L + (L % 2)), dtype=torch.float32, device=self.device)
            padded[:, :L] = wave
            wave = padded[:, :L]
        if lengths is not None:
            lengths = lengths.to(device=self.device, dtype=torch.int32).contiguous()
            if lengths.shape != (wave.shape[0],):
                raise ValueError("lengths must have one entry per clip")
        return wave, lengths

    @staticmethod
    def _row_stride(wave: torch.Tensor) -> int:
        return wave.stride(0) if wave.shape[0] > 1 else wave.shape[1]

    # ------------------------------------------------------------------ 8f-1: load_audio's device part, batched
    def load_audio(self, wave: torch.Tensor, orig_sample_rate: int, sample_rate: int = SAMPLE_RATE,
                   cut_time_seconds: float = 10, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``load_audio`` after the file decode (``utilityFunctions.py:109-120``): pad / cut to
        ``int(cut_time_seconds * orig_sample_rate)`` samples, ``torchaudio.functional.resample`` (defaults) to
        ``sample_rate``, mean of the two channels of a stereo clip.  ``wave`` is ``(B, C, L)`` (or ``(C, L)``),
        C in {1, 2}; ``lengths`` optionally gives the valid samples of each clip.  Returns ``(B, L')`` float32."""
        squeeze = wave.ndim == 2
        if squeeze:
            wave = wave.unsqueeze(0)
        if wave.ndim != 3:
            raise ValueError(f"wave must be (B, C, L) or (C, L), got {tuple(wave.shape)}")
        wave = wave.to(device=self.device, dtype=torch.float32).contiguous()
        B, C, L = wave.shape
        if lengths is not None:
            lengths = lengths.to(device=self.device, dtype=torch.int32).contiguous()
        key = (int(orig_sample_rate), int(sample_rate))
        r = self._resamplers.get(key)
        if r is None:
            r = ctypes.c_void_p()
            _lib.check(self.lib.ast_resampler_create(key[0], key[1], self.device.index, ctypes.byref(r)))
            self._resamplers[key] = r
        cut = int(cut_time_seconds * orig_sample_rate)
        n_out = int(self.lib.ast_resample_length(cut, key[0], key[1]))
        out = torch.empty((B, n_out), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ast_load_audio_forward(r, _ptr(wave), _ptr(lengths), B, C, L, cut, _ptr(out), n_out,
                                                       _stream_ptr(self.device)))
        return out[0:1] if squeeze else out

    # ------------------------------------------------------------------ synthetic dataset-scale input (configs[3])
    def synth_clips(self, n_clips: int, first_clip_id: int = 0, violin_from_id: int = 1 << 62,
                    n_samples: int = 220500, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``(n_clips, n_samples)`` float32 on the device: clip ``first_clip_id + b`` of the counter-based synthetic
        dataset (``ast_synth_clips``: piano-like below ``violin_from_id``, violin-like from there on).  A pure function of
        the clip id, so any rank regenerates any shard bit for bit."""
        if out is None:
            out = torch.empty((int(n_clips), int(n_samples)), dtype=torch.float32, device=self.device)
        elif out.ndim != 2 or out.shape[0] < n_clips or out.shape[1] < n_samples or out.stride(1) != 1 or out.device != self.device:
            raise ValueError("out must be a (>= n_clips, >= n_samples) float32 row-major tensor on the plan's device")
        done = 0
        with torch.cuda.device(self.device):
            while done < n_clips:  # the launch grid holds 65 535 clips
                n = min(32768, n_clips - done)
                _lib.check(self.lib.ast_synth_clips(_ptr(out[done:]), out.stride(0), n, int(n_samples),
                                                    int(first_clip_id) + done, int(violin_from_id), _stream_ptr(self.device)))
                done += n
        return out[:n_clips, :n_samples]

    # ------------------------------------------------------------------ a1: get_STFT, batched
    def stft(self, wave: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``(B, L)`` -> ``(B, 2, T, 513)``, ``T = 1 + L // 256`` (``utilityFunctions.py:12-37``)."""
        wave, lengths = self._wave(wave, lengths)
        B, L = wave.shape
        T = num_frames(L)
        out = torch.empty((B, 2, T, F_STFT), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ast_stft_forward(self._plan, _ptr(wave), _ptr(lengths), B, L, self._row_stride(wave),
                                                 _ptr(out), T, _stream_ptr(self.device)))
        return out

    # ------------------------------------------------------------------ a2: get_CQT, batched
    def cqt(self, wave: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``(B, L)`` -> ``(B, 2, T, 84)`` (``utilityFunctions.py:39-60``: ``librosa.cqt``)."""
        wave, lengths = self._wave(wave, lengths)
        B, L = wave.shape
        T = num_frames(L)
        out = torch.empty((B, 2, T, F_CQT), dtype=torch.float32, device=self.device)
        nbytes = self.lib.ast_workspace_bytes(self._plan, B, L)
        ws = self._workspace(nbytes)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ast_cqt_forward(self._plan, _ptr(wave), _ptr(lengths), B, L, self._row_stride(wave),
                                                _ptr(ws), nbytes, _ptr(out), T, _stream_ptr(self.device)))
        return out

    # ------------------------------------------------------------------ a1-a6 fused
    def features(self, wave: torch.Tensor, lengths: Optional[torch.Tensor] = None, mean: Optional[torch.Tensor] = None,
                 std: Optional[torch.Tensor] = None, eps: float = 1e-8, layout: str = "sections",
                 out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
        """STFT + CQT + normalise + concat (+ section cut) for a batch (``dataloader.py:100-112``).

        ``mean`` / ``std``: ``(2, 597)`` shared, ``(B, 2, 597)`` per clip, or ``None`` (raw features).
        Returns ``(features, counts)``: ``(B, S, 2, 287, 597)`` + sections per clip, or
        ``(B, 2, T, 597)`` + frames per clip."""
        wave, lengths = self._wave(wave, lengths)
        B, L = wave.shape
        T = num_frames(L)
        lay = _LAYOUTS[layout]
        if lay == LAYOUT_SECTIONS:
            dim1 = num_sections(T, self.window_size, self.overlap_frames)
            if dim1 == 0:
                raise RuntimeError(f"{T} frames give no section of {self.window_size} frames "
                                   "(torch.stack of an empty list in the reference, utilityFunctions.py:263)")
            shape = (B, dim1, 2, self.window_size, F_TOTAL)
        else:
            dim1 = T
            shape = (B, 2, T, F_TOTAL)
        per_clip = 0
        if (mean is None) != (std is None):
            raise ValueError("mean and std must both be given or both be None")
        if mean is not None:
            mean = mean.to(device=self.device, dtype=torch.float32).contiguous()
            std = std.to(device=self.device, dtype=torch.float32).contiguous()
            if mean.shape != std.shape:
                raise ValueError("mean and std shapes differ")
            if tuple(mean.shape) == (B, 2, F_TOTAL) and mean.ndim == 3:
                per_clip = 1
            elif tuple(mean.shape) != (2, F_TOTAL):
                raise ValueError(f"stats must be (2, {F_TOTAL}) or (B, 2, {F_TOTAL}), got {tuple(mean.shape)}")
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=self.device)
        elif tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != self.device:
            raise ValueError(f"out must be a contiguous float32 {shape} tensor on {self.device}")
        counts = torch.empty((B,), dtype=torch.int32, device=self.device)
        nbytes = self.lib.ast_workspace_bytes(self._plan, B, L)
        if workspace is None:
            ws = self._workspace(nbytes)
        else:  # caller-owned scratch (needed when several streams are in flight at once)
            ws = workspace
            if ws.numel() * ws.element_size() < nbytes or ws.device != self.device:
                raise ValueError(f"workspace needs {nbytes} bytes on {self.device}")
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ast_features_forward(
                self._plan, _ptr(wave), _ptr(lengths), B, L, self._row_stride(wave), _ptr(mean), _ptr(std), per_clip,
                float(eps), _ptr(ws), nbytes, _ptr(out), dim1, lay, _ptr(counts), _stream_ptr(self.device)))
        return out, counts

    def workspace_bytes(self, batch: int, n_samples: int) -> int:
        return int(self.lib.ast_workspace_bytes(self._plan, int(batch), int(n_samples)))

    def features_host(self, host_wave: torch.Tensor, host_out: torch.Tensor, mean: Optional[torch.Tensor] = None,
                      std: Optional[torch.Tensor] = None, eps: float = 1e-8, chunk: int = 16, n_streams: int = 3) -> torch.Tensor:
        """The host-buffer form of :meth:`features` (sections layout): ``host_wave (B, L)`` float32 in (pinned) host
        memory -> ``host_out (B, S, 2, 287, 597)`` float32 in (pinned) host memory.  Clips are cut into chunks; the
        H2D copy of chunk k + 1, the kernels of chunk k and the D2H copy of chunk k - 1 overlap (PCIe is full duplex).

        ALL kernels run on ONE compute stream (the caller's current stream); only the copies run on two side streams,
        ordered by events, over ``n_streams`` rotating device buffers.  The feature call's tensor-core kernels are
        persistent grids whose CTAs wait on counters written by other CTAs of the same call (``csrc/decimate_tc.cu``,
        ``csrc/cqt_tc.cu``): two such calls in flight on different streams could each be partly resident and starve one
        another, so at most one feature call per device is ever in flight (``include/ast_frontend.h``).  The kernels
        fill the GPU on their own, so nothing is lost.  Returns ``host_out`` after everything has landed."""
        if host_wave.ndim != 2 or host_wave.dtype != torch.float32 or host_wave.is_cuda:
            raise ValueError("host_wave must be a (B, L) float32 host tensor")
        B, L = host_wave.shape
        S = num_sections(num_frames(L), self.window_size, self.overlap_frames)
        shape = (B, S, 2, self.window_size, F_TOTAL)
        if tuple(host_out.shape) != shape or host_out.dtype != torch.float32 or host_out.is_cuda or not host_out.is_contiguous():
            raise ValueError(f"host_out must be a contiguous float32 host tensor of shape {shape}")
        chunk = max(1, min(int(chunk), B))
        if mean is not None:
            mean = mean.to(device=self.device, dtype=torch.float32).contiguous()
            std = std.to(device=self.device, dtype=torch.float32).contiguous()
        n_slots = max(2, int(n_streams))
        key = (chunk, L, S, n_slots)
        pipe = getattr(self, "_pipe", None)
        if pipe is None or pipe["key"] != key:
            with torch.cuda.device(self.device):
                pipe = {"key": key,
                        "h2d": torch.cuda.Stream(device=self.device), "d2h": torch.cuda.Stream(device=self.device),
                        # one scratch buffer: the compute stream serialises the feature calls that use it
                        "ws": torch.empty(self.workspace_bytes(chunk, L), dtype=torch.uint8, device=self.device),
                        "slots": [{
                            "x": torch.empty((chunk, L + (L % 2)), dtype=torch.float32, device=self.device),
                            "y": torch.empty((chunk,) + shape[1:], dtype=torch.float32, device=self.device),
                            "x_free": None,   # event: the kernels that read x have finished
                            "y_free": None,   # event: the D2H copy that read y has finished
                        } for _ in range(n_slots)]}
            self._pipe = pipe
        cur = torch.cuda.current_stream(self.device)
        h2d, d2h = pipe["h2d"], pipe["d2h"]
        h2d.wait_stream(cur)
        d2h.wait_stream(cur)
        for slot in pipe["slots"]:
            slot["x_free"] = slot["y_free"] = None
        for i, lo in enumerate(range(0, B, chunk)):
            hi = min(lo + chunk, B)
            n = hi - lo
            slot = pipe["slots"][i % n_slots]
            x = slot["x"][:n, :L]
            with torch.cuda.stream(h2d):
                if slot["x_free"] is not None:
                    h2d.wait_event(slot["x_free"])
                x.copy_(host_wave[lo:hi], non_blocking=True)
                x_ready = h2d.record_event()
            cur.wait_event(x_ready)
            if slot["y_free"] is not None:
                cur.wait_event(slot["y_free"])
            m, sd = mean, std
            if mean is not None and mean.ndim == 3:
                m, sd = mean[lo:hi], std[lo:hi]
            y, _ = self.features(x, mean=m, std=sd, eps=eps, layout="sections", out=slot["y"][:n], workspace=pipe["ws"])
            y_ready = cur.record_event()
            slot["x_free"] = y_ready
            with torch.cuda.stream(d2h):
                d2h.wait_event(y_ready)
                host_out[lo:hi].copy_(y, non_blocking=True)
                slot["y_free"] = d2h.record_event()
        cur.wait_stream(d2h)
        cur.wait_stream(h2d)
        cur.synchronize()
        return host_out

    # ------------------------------------------------------------------ a7-a9 fused
    def istft(self, spec: torch.Tensor, layout: str = "flat", overlap: Optional[int] = None,
              original_size: int = 0) -> torch.Tensor:
        """``(B, 2, T, F>=513)`` or ``(B, S, 2, 287, F>=513)`` -> ``(B, 256 * (T' - 1))``
        (``sections2spectrogram`` + ``inverse_STFT``, ``utilityFunctions.py:265-283``, ``:62-82``)."""
        lay = _LAYOUTS[layout]
        spec = spec.to(device=self.device, dtype=torch.float32).contiguous()
        overlap = self.overlap_frames if overlap is None else int(overlap)
        if lay == LAYOUT_FLAT:
            if spec.ndim != 4 or spec.shape[1] != 2:
                raise ValueError(f"flat spectrogram must be (B, 2, T, F), got {tuple(spec.shape)}")
            B, _, dim1, f_in = spec.shape
            n_frames = dim1
        else:
            if spec.ndim != 5 or spec.shape[2] != 2 or spec.shape[3] != self.window_size:
                raise ValueError(f"sections must be (B, S, 2, {self.window_size}, F), got {tuple(spec.shape)}")
            B, dim1, _, _, f_in = spec.shape
            n_frames = (self.window_size - overlap) * (dim1 - 1) + self.window_size
        if original_size and original_size > 0:
            n_frames = min(n_frames, int(original_size))
        n_out = max(0, HOP * (n_frames - 1))
        out = torch.empty((B, n_out), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ast_istft_forward(self._plan, _ptr(spec), B, dim1, f_in, lay, overlap, int(original_size),
                                                  _ptr(out), max(n_out, 1), _stream_ptr(self.device)))
        return out

    # ------------------------------------------------------------------ small operators
    def normalize(self, x: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        mean = mean.to(device=self.device, dtype=torch.float32).contiguous()
        std = std.to(device=self.device, dtype=torch.float32).contiguous()
        n_ch, n_time, n_freq = x.shape
        if tuple(mean.shape) != (n_ch, n_freq) or tuple(std.shape) != (n_ch, n_freq):
            raise ValueError("mean / std must be (channels, freq)")
        out = torch.empty_like(x)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ast_normalize(_ptr(x), _ptr(mean), _ptr(std), float(eps), n_ch, n_time, n_freq, _ptr(out),
                                              _stream_ptr(self.device)))
        return out

    def concat(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        a = a.to(device=self.device, dtype=torch.float32).contiguous()
        b = b.to(device=self.device, dtype=torch.float32).contiguous()
        n_ch, n_time, f1 = a.shape
        f2 = b.shape[2]
        out = torch.empty((n_ch, n_time, f1 + f2), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ast_concat(_ptr(a), _ptr(b), n_ch, n_time, f1, f2, _ptr(out), _stream_ptr(self.device)))
        return out

    def overlap_windows(self, spec: torch.Tensor, window_size: int, overlap_frames: int) -> torch.Tensor:
        spec = spec.to(device=self.device, dtype=torch.float32).contiguous()
        n_ch, n_time, n_freq = spec.shape
        n_sec = num_sections(n_time, window_size, overlap_frames)
        if n_sec == 0:
            raise RuntimeError("stack expects a non-empty TensorList")  # torch.stack([]) in the reference
        out = torch.empty((n_sec, n_ch, window_size, n_freq), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ast_overlap_windows(_ptr(spec), n_ch, n_time, n_freq, int(window_size), int(overlap_frames),
                                                    _ptr(out), n_sec, _stream_ptr(self.device)))
        return out

    def sections_merge(self, sections: torch.Tensor, original_size: int, overlap: int) -> torch.Tensor:
        sections = sections.to(device=self.device, dtype=torch.float32).contiguous()
        n_sec, n_ch, wind, n_freq = sections.shape
        n_time = (wind - overlap) * (n_sec - 1) + wind
        t_out = max(0, min(n_time, int(original_size)))
        out = torch.empty((n_ch, t_out, n_freq), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ast_sections_merge(_ptr(sections), n_sec, n_ch, wind, n_freq, int(overlap), t_out, _ptr(out),
                                                   _stream_ptr(self.device)))
        return out

    # ------------------------------------------------------------------ a10: dataset statistics
    def new_stats_accumulator(self, n_groups: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
        """``acc (G, 2, 2, 597)`` float64 (sum of clip means, sum of clip variances) and ``counts (G,)``."""
        acc = torch.zeros((n_groups, 2, 2, F_TOTAL), dtype=torch.float64, device=self.device)
        counts = torch.zeros((n_groups,), dtype=torch.float64, device=self.device)
        return acc, counts

    def _check_accumulator(self, acc: torch.Tensor, counts: torch.Tensor, f_dim: int = F_TOTAL) -> None:
        # the kernels add into these buffers through raw pointers: a host tensor or another dtype would be a wild write
        if (acc.ndim != 4 or tuple(acc.shape[1:]) != (2, 2, f_dim) or acc.dtype != torch.float64 or acc.device != self.device
                or not acc.is_contiguous()):
            raise ValueError(f"acc must be a contiguous float64 (G, 2, 2, {f_dim}) tensor on {self.device} "
                             "(new_stats_accumulator)")
        if (tuple(counts.shape) != (acc.shape[0],) or counts.dtype != torch.float64 or counts.device != self.device
                or not counts.is_contiguous()):
            raise ValueError(f"counts must be a contiguous float64 ({acc.shape[0]},) tensor on {self.device}")

    def stats_accumulate(self, wave: torch.Tensor, acc: torch.Tensor, counts: torch.Tensor,
                         lengths: Optional[torch.Tensor] = None, group_ids: Optional[torch.Tensor] = None) -> None:
        """Adds the per-clip per-bin mean / unbiased variance of the raw features of ``wave`` into ``acc``
        (``compute_separated_stats.py:16-43``)."""
        self._check_accumulator(acc, counts)
        wave, lengths = self._wave(wave, lengths)
        B, L = wave.shape
        if group_ids is not None:
            group_ids = group_ids.to(device=self.device, dtype=torch.int32).contiguous()
            if group_ids.shape != (B,):
                raise ValueError("group_ids must have one entry per clip")
        nbytes = self.lib.ast_stats_workspace_bytes(self._plan, B, L)
        ws = self._workspace(nbytes)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ast_stats_accumulate(self._plan, _ptr(wave), _ptr(lengths), _ptr(group_ids), B, L,
                                                     self._row_stride(wave), _ptr(ws), nbytes, acc.shape[0], _ptr(acc),
                                                     _ptr(counts), _stream_ptr(self.device)))

    def stats_accumulate_features(self, feats: torch.Tensor, acc: torch.Tensor, counts: torch.Tensor,
                                  n_frames: Optional[torch.Tensor] = None, group_ids: Optional[torch.Tensor] = None) -> None:
        feats = feats.to(device=self.device, dtype=torch.float32).contiguous()
        if feats.ndim != 4 or feats.shape[1] != 2:
            raise ValueError(f"feats must be (B, 2, T, F), got {tuple(feats.shape)}")
        B, _, t_dim, f_dim = feats.shape
        self._check_accumulator(acc, counts, f_dim)
        if n_frames is not None:
            n_frames = n_frames.to(device=self.device, dtype=torch.int32).contiguous()
        if group_ids is not None:
            group_ids = group_ids.to(device=self.device, dtype=torch.int32).contiguous()
        nbytes = 8 * 4 * f_dim * B
        ws = self._workspace(nbytes)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ast_stats_accumulate_features(self._plan, _ptr(feats), _ptr(n_frames), _ptr(group_ids), B,
                                                              t_dim, f_dim, _ptr(ws), nbytes, acc.shape[0], _ptr(acc),
                                                              _ptr(counts), _stream_ptr(self.device)))


_default: dict = {}


def default_frontend(device=None, window_size: int = WINDOW_SIZE, overlap_frames: int = OVERLAP_FRAMES) -> FrontEnd:
    """Process-wide cache of plans keyed by (device, window, overlap)."""
    if not torch.cuda.is_available():
        raise RuntimeError("audio-style-transfer_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        dev = torch.device("cuda", torch.cuda.current_device())
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    key = (dev.index, int(window_size), int(overlap_frames))
    fe = _default.get(key)
    if fe is None:
        fe = _default[key] = FrontEnd(dev, window_size, overlap_frames)
    return fe
