"""Vehicle constants of the path (values of vmvo/constants.py:3-7 in the reference)."""

WHEEL_BASE = 2.83972  # m
STEERING_RATIO = 13.27  # steering-wheel angle / road-wheel angle
MAX_STEER = 460.0  # degrees, steering wheel
MAX_ACCEL = 10  # m/s^2
MAX_STEER_RATE = 100.0  # degrees/s, steering wheel
