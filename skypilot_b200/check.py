"""Which clouds may be used (stand-in for sky/check.py:406-450).

The reference keeps the enabled-cloud list in its state database and the tests
monkeypatch `get_cached_enabled_clouds_or_refresh`; here the list is derived
from the loaded catalog (every cloud with a `vms.csv`) unless set explicitly.
"""
from typing import List, Optional, Sequence

from skypilot_b200 import exceptions

_enabled: Optional[List[str]] = None


def set_enabled_clouds(names: Optional[Sequence[str]]) -> None:
    global _enabled
    _enabled = None if names is None else [n.lower() for n in names]


def get_cached_enabled_clouds_or_refresh(capability=None,
                                         raise_if_no_cloud_access: bool = False):
    del capability
    from skypilot_b200 import catalog  # pylint: disable=import-outside-toplevel
    from skypilot_b200.utils import registry  # pylint: disable=import-outside-toplevel
    names = _enabled
    if names is None:
        store = catalog.get_store(required=False)
        names = [t.name for t in store.clouds] if store is not None else []
    clouds = []
    for name in names:
        if name in registry.CLOUD_REGISTRY or name in getattr(
                registry.CLOUD_REGISTRY, '_aliases', {}):
            clouds.append(registry.CLOUD_REGISTRY.from_str(name))
    if not clouds and raise_if_no_cloud_access:
        raise exceptions.NoCloudAccessError(
            'Cloud access is not set up. Load a catalog '
            '(skypilot_b200.catalog.load_*) or run `sky check`.')
    return clouds


def check_capability(*args, **kwargs) -> None:
    del args, kwargs
