"""B200-native (sm_100a) drop-in for the generator hot path of
OpheliaMiralles/wind-downscaling-gan (`src/downscaling`).  See DESIGN.md."""
from .api import *  # noqa: F401,F403  (mirrors downscaling/__init__.py:3)
