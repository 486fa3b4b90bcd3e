"""tinynerf_b200 -- B200-native (sm_100a) packed-ray render/train hot path of loicmagne/tinynerf.

    from tinynerf_b200.core import RayProvider, OccupancyGrid, NerfRenderer, NerfWeights, ...
    from tinynerf_b200.models import KPlanesFeatureField, CobafaFeatureField, VanillaOpacityDecoder, ...
    from tinynerf_b200 import _cuda   # compute_weights_fwd / compute_weights_bwd

Everything runs on hand-written CUDA kernels behind the C ABI of include/tinynerf_b200.h
(`libtinynerf_b200.so`, built by `python -m tinynerf_b200.build`).  There is no CPU fallback.
"""
__version__ = "0.1.0"
