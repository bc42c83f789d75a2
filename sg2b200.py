"""Importable alias for the package directory `speech-to-image-translation-without-text_b200/`
(a hyphenated directory name cannot be written in an import statement)."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "speech-to-image-translation-without-text_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
