"""gonzag_b200: the gridded -> along-track mapping + interpolation path of brodeau/gonzag,
rebuilt for NVIDIA B200 (sm_100a).  Same Python surface as `gonzag.bilin_mapping` and
`gonzag.mod2sat`; the work is done by hand-written CUDA kernels behind a C ABI
(include/gonzag_b200.h, gonzag_b200/libgonzag_b200.so).  No CPU fallback."""
from .config import *                                   # noqa: F401,F403
from ._lib import GonzagB200Error, device_count        # noqa: F401
from .context import GridContext                        # noqa: F401
from .bilin_mapping import (BilinTrack, NearestPoint, Iquadran, IDSourceMesh, AlfaBeta, WeightBL,   # noqa: F401
                            Heading, IQH, surrounding_j_i, Iquadran2SrcMesh, release_cached_context)
from .mod2sat import Model2SatTrack, RecordWindow       # noqa: F401
from .sharded import ShardedModel2SatTrack, replicate_grid   # noqa: F401
from .cache import save_mapping, load_mapping           # noqa: F401
from .utils import (GridAngle, SearchBoxSize, degE_to_degWE, GridResolution, IsEastWestPeriodic,       # noqa: F401
                    IsGlobalLongitudeWise, scan_idx, GetEpochTimeOverlap, EpochT2Str, ModGrid, SatTrack, MsgExit,
                    Haversine, Haversine_sclr, RadiusEarth, find_j_i_min, find_j_i_max, chck4f)
from .sources import ModelArrays, TrackArrays           # noqa: F401
from .nc3io import NetCDF3Source                        # noqa: F401
from .ncout import SaveTimeSeries, Save2Dfield          # noqa: F401

__all__ = ["BilinTrack", "NearestPoint", "Iquadran", "IDSourceMesh", "AlfaBeta", "WeightBL", "Heading", "IQH",
           "surrounding_j_i", "Iquadran2SrcMesh", "Haversine", "Haversine_sclr", "RadiusEarth", "find_j_i_min",
           "find_j_i_max", "chck4f", "MsgExit",
           "Model2SatTrack", "ShardedModel2SatTrack", "RecordWindow", "replicate_grid", "save_mapping", "load_mapping", "GridContext", "GridAngle", "SearchBoxSize",
           "degE_to_degWE", "GonzagB200Error", "device_count", "GridResolution", "IsEastWestPeriodic",
           "IsGlobalLongitudeWise", "scan_idx", "GetEpochTimeOverlap", "EpochT2Str", "ModGrid", "SatTrack",
           "ModelArrays", "TrackArrays", "NetCDF3Source", "SaveTimeSeries", "Save2Dfield"]
