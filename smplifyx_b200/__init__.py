"""Importable alias for the ``smplify-x-partial_b200/`` package directory.

The product package lives in ``smplify-x-partial_b200/`` (the name the build
contract asks for).  A hyphenated directory is not a valid Python identifier,
so this one-file package re-points its ``__path__`` at that directory: every
``import smplifyx_b200.<module>`` resolves to a file under
``smplify-x-partial_b200/``.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
_real = _os.path.join(_os.path.dirname(_here), 'smplify-x-partial_b200')
if not _os.path.isdir(_real):  # pragma: no cover
    raise ImportError('smplify-x-partial_b200/ directory is missing next to smplifyx_b200/')
__path__ = [_real]

with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
