"""Shaders -- host-side mirror of the reference's shader.py.

PhongShader / DepthMapShader select which shading the render kernels apply
(`shader_id`).  `PhongShader(specular=False)` is the orbit_experiments variant
whose specular term is commented out (orbit_experiments/shader.py:45,48).
`shade()` is the dense per-shape evaluation kept for API compatibility (torch
ops); Scene.build() never calls it.
"""
import torch

from . import _native as nat
from .util import broadcasted_switch


class Shader(object):
    shader_id = None
    maxDepth = 1.0


class DepthMapShader(Shader):
    """shader.py:9-20"""
    shader_id = nat.SHADER_DEPTH

    def __init__(self, maxDepth):
        self.maxDepth = float(maxDepth)

    def shade(self, shape, lights, camera):
        distance = shape.distance(camera.rays)
        scaled = (distance - 0) / (self.maxDepth - 0)
        return (1 - scaled).unsqueeze(-1) * torch.ones(3, dtype=scaled.dtype, device=scaled.device)


class PhongShader(Shader):
    """shader.py:23-53"""

    def __init__(self, specular=True):
        self.specular = bool(specular)
        self.shader_id = nat.SHADER_PHONG if specular else nat.SHADER_PHONG_NOSPEC

    def shade(self, shape, lights, camera):
        light = lights[0]                                   # shader.py:33
        material = shape.material
        normals = shape.normals(camera.rays)
        Lh = light.normed_dir().to(normals.device)
        ndl = (normals * (-Lh)).sum(2)
        phong = material.ka + material.kd * ndl
        if self.specular:
            rm = 2.0 * ndl.unsqueeze(-1) * normals + Lh
            look = torch.as_tensor(camera.look_at, dtype=torch.float32, device=normals.device)
            phong = phong + material.ks * ((rm * look).sum(2) ** material.shininess)
        colorized = phong.unsqueeze(-1) * material.color.to(normals.device) * light.intensity.to(normals.device)
        clipped = torch.clamp(colorized, 0, 1)
        distances = shape.distance(camera.rays)
        return broadcasted_switch(torch.isinf(distances), [0., 0., 0.], clipped)
