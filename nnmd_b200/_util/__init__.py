from ._logger import Logger

__all__ = ["Logger"]
