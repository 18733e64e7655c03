"""CPU oracle for the spectral front-end.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker / the CPU timing arm.
The product path (``audio-style-transfer_b200``) never imports it and fails
loudly when its CUDA library is missing.

Parity status:

* STFT / iSTFT / section cut / section merge / normalise / concat / collate /
  dataset statistics are **pinned**: ``oracle/make_golden.py`` imports the
  unmodified reference (``/root/reference/utilityFunctions.py``,
  ``dataloader.py``, with ``librosa`` / ``matplotlib`` stubbed because they are
  not installed) in the build container, runs it on seeded inputs and commits
  the outputs under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks
  this restatement against those vectors.
* CQT is **parity unpinned**: the arithmetic lives in ``librosa`` (version not
  pinned by the reference, README.md:160-168) and ``soxr``; neither is vendored
  under ``/root/reference`` nor installable here.  ``oracle/cqt.py`` restates
  the documented librosa >= 0.10 algorithm and a soxr-HQ-like decimator; the
  reference's own tests pin only the output shape ``(2, 862, 84)``
  (test_correctness.ipynb cell 3).
"""
