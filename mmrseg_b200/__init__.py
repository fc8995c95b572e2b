"""Importable alias of the package directory `mmr_semantic-segmentation_v1_b200/`.

The directory name demanded by the repo layout contains a hyphen, which Python cannot
import; this stub points the `mmrseg_b200` package at that directory so that
`import mmrseg_b200.plan` loads `mmr_semantic-segmentation_v1_b200/plan.py`.
"""
import os as _os

_real = _os.path.normpath(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..",
                                        "mmr_semantic-segmentation_v1_b200"))
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
