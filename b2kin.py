"""Import alias: the package directory `robotics-toolbox-python_b200/` is not a valid Python
identifier, so `import b2kin` loads it through importlib and re-exports it.

    import b2kin as rtb
    panda = rtb.models.Panda()
    T, J = panda.fkine_jacob0(Q)          # Q: (N, 7) numpy array or CUDA torch tensor
"""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("robotics-toolbox-python_b200")
sys.modules[__name__] = _pkg
