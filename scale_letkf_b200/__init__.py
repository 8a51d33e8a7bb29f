"""scale_letkf_b200 -- B200-native (sm_100a, fp64) LETKF analysis hot path behind the
reference's letkf_core / das_letkf interface.  See DESIGN.md."""
from . import capi, config, synth  # noqa: F401
from .api import LETKF, LetkfError  # noqa: F401
from .config import default_config, resolve_config  # noqa: F401
