"""Import alias: the product package lives in the (non-importable) directory
``multimodal-gesture-recognition-with-lstms-and-ctc_b200/``; ``import mgr_b200`` resolves to it."""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
__path__ = [_os.path.join(_os.path.dirname(_here), "multimodal-gesture-recognition-with-lstms-and-ctc_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
