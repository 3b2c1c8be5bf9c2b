"""Chunk-size helpers kept for API parity with `src/jaxhps/_device_config.py:37-90`.

The reference needs these because its assembled leaf operators (23.9 MB per leaf at p=12) do not fit
an 80 GB device; here the leaf stage bounds its own scratch (13.8 MB per leaf, chunked internally), so
the numbers only matter to callers that drive the stages chunk by chunk themselves (subtree
recomputation)."""
from __future__ import annotations


def local_solve_chunksize_2D(p: int, dtype) -> int:
    """Always 4**7, as in the reference (its earlier branches are overwritten)."""
    return 4**7


def local_solve_chunksize_3D(p: int, dtype) -> int:
    """Leaves per device batch the reference uses on an 80 GB GPU (`_device_config.py:65-90`)."""
    if p <= 8:
        return 2_000
    if p <= 10:
        return 500
    if p <= 12:
        return 100
    return 20
