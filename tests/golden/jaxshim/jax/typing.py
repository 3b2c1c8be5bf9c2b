import numpy as _np

DTypeLike = _np.dtype
ArrayLike = object
