"""B200-native vehicle-model-constrained window search (the optimize_trajectory_v2 path of
AdityaNG/VehicleModelVisualOdometry), behind the reference's Python signatures.

Importing the package does not need a GPU; every operator does, and raises otherwise.
"""
from .constants import MAX_ACCEL, MAX_STEER, MAX_STEER_RATE, STEERING_RATIO, WHEEL_BASE  # noqa: F401
from .schema import State, Trajectory, states_list_to_trajectory  # noqa: F401
from .bicycle_model import BicycleModel, rollout_batch  # noqa: F401
from .search import (DrivePipeline, DriveSet, DriveStream, SearchConfig, SearchOutput, WindowPlan,  # noqa: F401
                     grid_search, hypothesis_steps, optimize_drives, plan_windows, write_back)
from .mpc import grid_run, mpc_run, sequence_cost, traverse_trajectory  # noqa: F401
from .optimize import DEFAULT_CFG, REFERENCE_CFG, VO_CFG, optimize_trajectory  # noqa: F401
from .dataset import (load_android_drive, load_android_drives_device, optimize_android_drives,  # noqa: F401
                      parse_csv_files, prepare_android_drives, read_csv)

__version__ = "0.1.0"
