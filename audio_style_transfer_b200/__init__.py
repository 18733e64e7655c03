"""Import alias: the package directory is ``audio-style-transfer_b200/`` (not a valid Python
identifier), so ``import audio_style_transfer_b200`` resolves to it through this shim."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "audio-style-transfer_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__, "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), __file__, "exec"))
del _f, _real, _os
