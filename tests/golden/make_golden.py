"""Extracts the reference's own golden artefacts into small fixtures.

Run once in the build container (needs /root/reference and PIL):
    python tests/golden/make_golden.py
Only DATA is copied (rendered images the reference ships), never sources.

  orbit_sample99.npz   /root/reference/orbit_experiments/orbit_dataset.npz arr_0[99]
                       (2 views, 64x64x3 uint8, rendered by the reference renderer via
                       orbit_experiments/planet_orbit.py:55-67) + the sphere centre
                       from orbit_target.npz (only sample 99's centre survives, a bug
                       at planet_orbit.py:67).
  orbit_samples_0_7.npz  arr_0[0:8] (centres unknown; circle constraint
                       x^2+y^2=81, z=32, planet_orbit.py:44-53) for the fit test.
  match_mirror_frame0.npy  /root/reference/output/0.jpg decoded (128x128x3 uint8):
                       first frame of match_mirror.py (root camera variant,
                       Phong with specular, 2 spheres + 1 square). JPEG-lossy.
  balls_15.npy         /root/reference/15.jpg decoded (32x32 uint8): the
                       test_balls.py training target (test_balls.py:17) -- sample 15 of
                       generate_data.py's depth dataset: two unit spheres, DepthMapShader(6.1),
                       written by util.draw -> scipy.misc.imsave (rescaled to 0..255 by its max).
  example_png.npy      /root/reference/example.png decoded (32x32 uint8): the test_1ball.py
                       training target (test_1ball.py:17), one unit sphere, same pipeline, lossless.
  orbit_samples_all.npz  arr_0[0:99] (all remaining samples of the orbit dataset; centres unknown,
                       circle constraint) for the full-dataset fit test.
"""
import os
import numpy as np
from PIL import Image

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))

d = np.load(os.path.join(REF, 'orbit_experiments/orbit_dataset.npz'))['arr_0']
t = np.load(os.path.join(REF, 'orbit_experiments/orbit_target.npz'))['arr_0']
np.savez_compressed(os.path.join(HERE, 'orbit_sample99.npz'), views=d[99], centre=t)
np.savez_compressed(os.path.join(HERE, 'orbit_samples_0_7.npz'), views=d[0:8])
np.save(os.path.join(HERE, 'match_mirror_frame0.npy'),
        np.asarray(Image.open(os.path.join(REF, 'output/0.jpg'))))
np.save(os.path.join(HERE, 'balls_15.npy'),
        np.asarray(Image.open(os.path.join(REF, '15.jpg'))))
np.save(os.path.join(HERE, 'example_png.npy'),
        np.asarray(Image.open(os.path.join(REF, 'example.png'))))
np.savez_compressed(os.path.join(HERE, 'orbit_samples_all.npz'), views=d[0:99])
print('ok')
