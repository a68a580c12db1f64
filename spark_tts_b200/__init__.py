"""Importable alias for the ``spark-tts_b200/`` source directory.

The repository layout names the package directory ``spark-tts_b200`` (hyphenated, not a
valid Python identifier).  This shim extends ``__path__`` so that
``import spark_tts_b200`` resolves every sub-module from that directory.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_src = _os.path.join(_os.path.dirname(_here), "spark-tts_b200")
__path__.insert(0, _src)

with open(_os.path.join(_src, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_src, "__init__.py"), "exec"))
