"""Importable name for the package that lives in `edsnet-efficient-dsnet-for-video-summarization_b200/`
(a directory name Python cannot import).  Submodules (`edsnet_b200.dsnet`, `edsnet_b200.plan`, ...) resolve
to the files of that directory; nothing is duplicated here."""
import os as _os

_REAL = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "edsnet-efficient-dsnet-for-video-summarization_b200")
__path__.insert(0, _REAL)
with open(_os.path.join(_REAL, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_REAL, "__init__.py"), "exec"), globals())
