"""Importable alias of the package directory ``mdgat-matcher_b200/`` (a hyphen cannot appear
in a Python import statement). All code lives in ../mdgat-matcher_b200/."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'mdgat-matcher_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
