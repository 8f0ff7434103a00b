"""reversible_raytracer_b200 -- B200-native (sm_100a) differentiable ray tracer.

Drop-in for the one hot path of lebek/reversible-raytracer (primary rays ->
ray/shape intersection -> nearest hit -> Phong / depth shading -> reverse pass to
scene-parameter gradients) behind the reference's Scene / shape / shader /
transform Python API.  See DESIGN.md and INTEGRATION.md.
"""
from . import _native  # noqa: F401

__version__ = '0.1.0'
