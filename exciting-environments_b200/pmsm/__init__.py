from .motor_parameters import MotorVariant  # noqa: F401
