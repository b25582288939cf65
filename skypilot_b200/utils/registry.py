"""Name -> class registry the clouds plug into (mirrors sky/utils/registry.py:80-117)."""
from typing import Callable, Dict, Generic, List, Optional, Type, TypeVar

T = TypeVar('T')


class _Registry(Generic[T], dict):
    """`@REGISTRY.register` class decorator; `from_str()` returns an instance."""

    def __init__(self, registry_name: str, exclude: Optional[set] = None):
        super().__init__()
        self._name = registry_name
        self._exclude = exclude or set()
        self._aliases: Dict[str, str] = {}

    def from_str(self, name: Optional[str]) -> Optional[T]:
        if name is None:
            return None
        key = name.lower()
        if key in self._exclude:
            return None
        key = self._aliases.get(key, key)
        if key not in self:
            raise ValueError(f'{self._name.capitalize()} {name!r} is not a '
                             f'valid {self._name} among {list(self.keys())}')
        return self[key]()

    def register(self, cls: Optional[Type[T]] = None, *,
                 aliases: Optional[List[str]] = None) -> Callable:

        def _do(klass):
            key = klass.__name__.lower()
            if key in self:
                raise ValueError(f'{self._name} {key} already registered')
            self[key] = klass
            for alias in aliases or []:
                self._aliases[alias.lower()] = key
            return klass

        if cls is not None:
            return _do(cls)
        return _do


CLOUD_REGISTRY: _Registry = _Registry('cloud', exclude={'dummycloud'})
