import numpy as _np
from .._core import wrap_fn as _w

inv = _w(_np.linalg.inv)
solve = _w(_np.linalg.solve)
cond = _w(_np.linalg.cond)
norm = _w(_np.linalg.norm)
