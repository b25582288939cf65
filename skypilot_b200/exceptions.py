"""Exception classes the optimizer path raises (same names as sky/exceptions.py)."""
from typing import List, Optional


class ResourcesUnavailableError(Exception):
    """No launchable resource satisfies a task (sky/exceptions.py)."""

    def __init__(self,
                 message: str,
                 no_failover: bool = False,
                 failover_history: Optional[List[Exception]] = None):
        super().__init__(message)
        self.no_failover = no_failover
        self.failover_history = failover_history or []


class NoCloudAccessError(Exception):
    """No enabled cloud (sky/exceptions.py)."""


class NotSupportedError(Exception):
    """A cloud does not implement a requested feature (sky/exceptions.py)."""


class OptimizerLimitError(NotSupportedError):
    """The problem is valid but exceeds a documented capacity of this
    optimizer (e.g. a densely connected general DAG under the TIME objective).
    Deliberately NOT a ResourcesUnavailableError: failover loops must not read
    it as "no capacity in this region"."""


class ResourcesMismatchError(Exception):
    """The accelerators cannot be attached to the instance type."""


class InvalidCloudConfigs(Exception):
    """Invalid cloud configuration."""
