import numpy as _np
from .._core import wrap_fn as _w

ifft = _w(_np.fft.ifft)
fft = _w(_np.fft.fft)
ifftshift = _w(_np.fft.ifftshift)
