"""Mirror of the reference's models/networks/correlation_layer.py (same names, arguments and defaults)."""
from . import ops


def compute_cost_volume(features1, features2, max_displacement=4):
    """Cost volume between features1 and features2 displaced by up to max_displacement in x and y.

    Replaces reference correlation_layer.py:7-40.  features: [b, c, h, w] CUDA fp32.
    Returns [b, (2*max_displacement+1)**2, h, w]; channel k = (dy+d)*(2d+1)+(dx+d); mean over c; zero padding.
    """
    return ops.cost_volume(features1, features2, max_displacement)


def normalize_features(feature_list, normalize=True, center=True, moments_across_channels=True,
                       moments_across_images=True):
    """Replaces reference correlation_layer.py:42-82: returns the list of normalised tensors; gradients flow through
    the statistics exactly as in the reference (they are not detached)."""
    return ops.normalize_features(feature_list, normalize=normalize, center=center,
                                  moments_across_channels=moments_across_channels,
                                  moments_across_images=moments_across_images)
