"""Host-side data formats either side of the hot path."""
from .io import read_ply, save_ply
from .dataset import MVRData, decompose_to_R_and_t

__all__ = ["read_ply", "save_ply", "MVRData", "decompose_to_R_and_t"]
