"""b200ssl -- B200-native semi-supervised loss-and-mixing path.

Importable as `b200ssl` (alias package at the repo root) because this directory's name carries a
hyphen.  Modules mirror the reference's namespaces so that

    import b200ssl.cowmix as cowmix
    import b200ssl.lovasz as lovasz
    import b200ssl.mean_teacher as mean_teacher
    import b200ssl.metrics as metrics

is the whole patch to train.py / losses.py.
"""
from . import _lib            # raises ImportError if the CUDA library has not been built
from . import cowmix, lovasz, mean_teacher, metrics, losses, utils, consistency, optim  # noqa: F401
from .step import LossPathStep  # noqa: F401

__version__ = "0.1.0"
