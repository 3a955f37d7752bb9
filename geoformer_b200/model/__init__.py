"""Mirror of the reference's ``model`` package surface (the three modules the matcher wrappers import,
``eval_tool/immatch/modules/geoformer.py:6-10`` / ``inference.py:6-10``).

Importable under two names:
* ``geoformer_b200.model`` — as part of this package;
* top-level ``model`` — the drop-in: a reference script puts ``<repo>/geoformer_b200`` in front of ``sys.path`` and its
  unchanged ``from model.full_model import GeoFormer`` resolves here.  In that case the parent package
  (``geoformer_b200``: engine, ops, the ctypes binding) is not importable by name yet, so its parent directory is
  appended to ``sys.path`` here.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

if _ilu.find_spec("geoformer_b200") is None:          # imported as top-level `model`
    _sys.path.append(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
