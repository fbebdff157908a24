from .params import CameraParams
from .projection import apply, project_to_image

__all__ = ["CameraParams", "apply", "project_to_image"]
