"""Vehicle constants of the path.

The five values the reference's model and optimizer read (vmvo/constants.py:3-7); the names are
part of the interface (``from vmvo.constants import WHEEL_BASE, ...`` in bicycle_model.py:6-12 and
utils/mpc.py:8-12), so they are kept.  Derived quantities used by the kernels and the docs sit
beside them.
"""
import math

# axle-to-axle distance L of the kinematic bicycle: theta' = theta + (v / L) tan(delta) dt
WHEEL_BASE = 2.83972            # [m]
# steering-wheel degrees per road-wheel degree: delta = radians(s) / STEERING_RATIO
STEERING_RATIO = 13.27
# lock-to-lock limit of the steering wheel, asserted per step (bicycle_model.py:48-50)
MAX_STEER = 460.0               # [deg, steering wheel]
# |dv / dt| limit asserted per step (bicycle_model.py:59-62); spans the V axis of the hypothesis grid
MAX_ACCEL = 10                  # [m/s^2]
# steering-rate limit (stored by the reference, bound disabled at mpc.py:96-104); spans the S axis
MAX_STEER_RATE = 100.0          # [deg/s, steering wheel]

# largest road-wheel angle the limits allow: 0.605 rad -- the tangent polynomial of the search's TL
# table (csrc/vmvo_device.cuh: tan_steer) is exact to 1 ulp up to 0.62 rad
MAX_ROAD_WHEEL_ANGLE = math.radians(MAX_STEER) / STEERING_RATIO
