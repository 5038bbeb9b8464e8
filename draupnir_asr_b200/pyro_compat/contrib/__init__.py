from . import easyguide  # noqa: F401
