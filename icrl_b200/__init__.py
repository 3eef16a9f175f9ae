"""Importable alias of the product package.

The product lives in ``image-captioning-through-rl_b200/`` (the directory name the project
mandates; not a valid Python identifier), so this one-file package points its ``__path__`` there:
``import icrl_b200.models`` loads ``image-captioning-through-rl_b200/models.py``.
"""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "image-captioning-through-rl_b200")
__path__ = [_REAL]
LIB_DIR = _REAL
