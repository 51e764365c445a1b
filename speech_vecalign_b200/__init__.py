"""Importable alias of the package directory ``speech-vecalign_b200/`` (a hyphen is not a valid
Python identifier): ``import speech_vecalign_b200`` executes that package's ``__init__`` with its
``__path__``, so ``speech_vecalign_b200.dp_utils`` etc. resolve to the files under
``speech-vecalign_b200/``.  No code lives here."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "speech-vecalign_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
