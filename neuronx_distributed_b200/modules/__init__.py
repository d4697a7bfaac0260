from . import qkv_linear, rms_norm  # noqa: F401
