from . import expert_mlps_mx, mx_torch, swizzle  # noqa: F401
