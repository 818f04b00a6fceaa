"""Constants and small containers shared with the reference."""
from collections import namedtuple

# feabas/constant.py:39-41
FFT_CONF_NONE = 0
FFT_CONF_STD = 1
FFT_CONF_MIRROR = 2

# feabas/constant.py:6-10 -- vertex states ("gears") of a mesh
MESH_GEAR_INITIAL = -1
MESH_GEAR_FIXED = 0
MESH_GEAR_MOVING = 1
MESH_GEAR_STAGING = 2

# feabas/constant.py:18-21 -- MeshRenderer.crop modes
RENDER_FULL = 3

# feabas/constant.py:27-31
ANNEAL_CONNECTED_RIGID = 2
ANNEAL_COPY_EXACT = 4

# feabas/constant.py:43-44, feabas/config.py:32
DEFAULT_RESOLUTION = 4.0
DEFAULT_THICKNESS = 30.0
DEFAULT_AVG_DEFORM = 0.05

# feabas/common.py:18
Match = namedtuple('Match', ('xy0', 'xy1', 'weight', 'strain'), defaults=(DEFAULT_AVG_DEFORM,))
