"""Alias package: `semi-supervised_semantic_segmentation_b200/` is not a valid Python identifier,
so `import b200ssl` loads that directory as the package `b200ssl`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "semi-supervised_semantic_segmentation_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
