"""Stand-in for ``colorlog`` (TEST-ONLY): ``flashy/logging.py:53-57`` builds a ``ColoredFormatter``
whose format string uses colour fields; here they expand to nothing."""
import logging

_FIELDS = ("log_color", "reset", "cyan", "blue", "red", "green", "yellow", "purple", "white", "black",
           "bold", "thin")


class ColoredFormatter(logging.Formatter):
    def __init__(self, fmt=None, datefmt=None, style="%", **_kwargs):
        super().__init__(fmt, datefmt, style)

    def format(self, record):
        for name in _FIELDS:
            if not hasattr(record, name):
                setattr(record, name, "")
        return super().format(record)
