"""studiosr_b200 -- B200-native (sm_100a) execution of StudioSR's super-resolution hot path behind
the unchanged `studiosr.models` API.  See DESIGN.md / INTEGRATION.md."""
from . import models  # noqa: F401
from ._lib import LIB_PATH, load  # noqa: F401

__version__ = "0.1.0"
