"""Exception types: Gymnasium's own when available (gymnasium/error.py:4-98), same-named stand-ins otherwise."""
try:
    from gymnasium.error import (AlreadyPendingCallError, ClosedEnvironmentError, DependencyNotInstalled, Error, NoAsyncCallError,
                                 ResetNeeded)
except ImportError:  # pragma: no cover

    class Error(Exception):
        pass

    class ResetNeeded(Error):
        pass

    class DependencyNotInstalled(Error):
        pass

    class AlreadyPendingCallError(Exception):
        def __init__(self, message: str, name: str):
            super().__init__(message)
            self.name = name

    class NoAsyncCallError(Exception):
        def __init__(self, message: str, name: str):
            super().__init__(message)
            self.name = name

    class ClosedEnvironmentError(Exception):
        pass


__all__ = ["Error", "ResetNeeded", "DependencyNotInstalled", "AlreadyPendingCallError", "NoAsyncCallError",
           "ClosedEnvironmentError"]
