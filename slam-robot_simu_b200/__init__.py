"""B200-native Monte-Carlo-localisation engine: the particle_filter.py step of SLAM-Robot_Simu.

The directory name carries a hyphen, so import it with
``importlib.import_module("slam-robot_simu_b200")`` or through the root-level alias module
``slam_robot_simu_b200``.  Importing the package loads libmcl_b200.so and raises if it is missing.
"""
from . import _lib
from ._lib import MclError, mcl_config, mcl_shard_plan, mcl_stats

_lib.load()  # fail loudly at import time when the CUDA library has not been built

from .engine import Filter, default_config, stats_to_dict  # noqa: E402
from .particle_filter import ParticleFilter, LazyParticles  # noqa: E402
from .sharded import LocalShardedFilter, ShardedParticleFilter  # noqa: E402
from .scan_sensor import ScanSensor  # noqa: E402
from .trajectory import TrajectoryLog  # noqa: E402

__all__ = ["Filter", "ParticleFilter", "LazyParticles", "ShardedParticleFilter", "LocalShardedFilter", "ScanSensor", "TrajectoryLog", "MclError", "mcl_config", "mcl_stats", "default_config",
           "stats_to_dict"]
