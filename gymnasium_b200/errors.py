"""Exception types: Gymnasium's own when available (gymnasium/error.py:4-98), same-named stand-ins otherwise."""
try:
    from gymnasium.error import DependencyNotInstalled, Error, ResetNeeded
except ImportError:  # pragma: no cover

    class Error(Exception):
        pass

    class ResetNeeded(Error):
        pass

    class DependencyNotInstalled(Error):
        pass


__all__ = ["Error", "ResetNeeded", "DependencyNotInstalled"]
