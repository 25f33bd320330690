"""Replaces the reference's ``miscc`` package for ``config`` and ``utils`` only.  When this directory is
put FIRST on ``sys.path`` it shadows the reference's ``miscc``; the modules it does
not replace (``miscc/datasets.py``, imported by the reference's main_pororo.py:23 and inference.py:26) must
still resolve, so every other ``miscc`` directory found on ``sys.path`` is appended to this package's
search path."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
for _p in list(sys.path):
    _cand = os.path.abspath(os.path.join(_p or ".", "miscc"))
    if _cand != _here and os.path.isdir(_cand) and _cand not in __path__:
        __path__.append(_cand)
