"""Exception type raised by the pair counters (mirrors halotools/custom_exceptions.py:8)."""

__all__ = ("HalotoolsError",)


class HalotoolsError(Exception):
    """Base class of all exceptions raised by this package (same name as the reference's)."""

    def __init__(self, message):
        super(HalotoolsError, self).__init__(message)
