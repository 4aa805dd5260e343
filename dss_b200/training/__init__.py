"""Regularisers next to the hot path: the projection / repulsion losses of DSS/training/losses.py on top
of the B200 K-NN."""
from .losses import BaseLoss, L1Loss, L2Loss, ProjectionLoss, RepulsionLoss, SurfaceLoss  # noqa: F401
