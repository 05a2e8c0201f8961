"""``src/quansino/utils/strings.py``: the header format of a Logger column."""

from quansino_b200.io.logger import get_auto_header_format

__all__ = ["get_auto_header_format"]
