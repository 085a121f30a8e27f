from . import nms_wrapper  # noqa: F401
