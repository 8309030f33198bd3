"""A tiny stand-in for the `tensorflow` names user `main.py` files touch.

The reference resolves normalisation functions by name from the user's `main.py`
(`code/utils/generate_model.py:24, 68, 77`) and the shipped examples call `tf.math.log` /
`tf.math.exp` inside them (`examples/Routenet/main.py:26-38`, `examples/Q-size/main.py:27-39`).
Normalisation is host-side work here (NumPy on the dataset arrays), so this shim maps those calls
onto NumPy.  `install()` registers it as `tensorflow` ONLY when the real package is absent.
"""

import sys
import types

import numpy as np


class _Math(types.SimpleNamespace):
    pass


math = _Math(log=np.log, exp=np.exp, sqrt=np.sqrt, log1p=np.log1p, abs=np.abs, pow=np.power,
             square=np.square, tanh=np.tanh, maximum=np.maximum, minimum=np.minimum,
             reduce_mean=np.mean, reduce_sum=np.sum, reduce_max=np.max, reduce_min=np.min)
float32 = np.float32
int64 = np.int64
log = np.log
exp = np.exp


def cast(x, dtype):
    return np.asarray(x).astype(dtype)


def install():
    """Make `import tensorflow as tf` resolve to this shim when TensorFlow is not installed."""
    try:
        import tensorflow  # noqa: F401
        return False
    except Exception:
        mod = sys.modules[__name__]
        sys.modules["tensorflow"] = mod
        return True
