"""Package version (reference ``_version.py``)."""
__version__ = "0.2.0"
