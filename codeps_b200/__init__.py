"""codeps_b200 -- B200-native photometric reprojection loss for CoDEPS."""
from .camera import CameraModel  # noqa: F401
