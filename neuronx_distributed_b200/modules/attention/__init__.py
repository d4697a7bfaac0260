from .utils import apply_rotary_polar_compatible, precompute_freqs_cis  # noqa: F401
