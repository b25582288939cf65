"""Import the UNMODIFIED reference (`/root/reference/sky`) offline, as an oracle.

Test infrastructure only. Nothing under `skypilot_b200/` may import this file.
It only works in the build container (the GPU box has no `/root/reference`);
its job is (1) to validate `oracle/*.py` (the portable restatement) and (2) to
generate the committed fixtures under `tests/golden/` (see `gen_golden.py`).

Recipe:
  * `colorama` / `prettytable` hand stubs (`stubs/`), MagicMock modules for
    the DB / auth dependencies the optimizer path never touches;
  * a scratch `$HOME` holding `~/.sky/catalogs/v8/<cloud>/vms.csv` written by
    hand -- a CSV without a `.meta/*.md5` sidecar counts as user-modified and
    is never re-downloaded (reference sky/catalog/common.py:71-84, :189-197);
  * `sky.check.get_cached_enabled_clouds_or_refresh` monkeypatched the way the
    reference's own test fixture does (tests/common_test_fixtures.py:175-236).
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from typing import Dict, Optional, Sequence
from unittest import mock

REFERENCE_ROOT = os.environ.get('SKYOPT_REFERENCE_ROOT', '/root/reference')
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'stubs')
_MOCKED = {
    'sqlalchemy', 'aiosqlite', 'alembic', 'passlib', 'psycopg2', 'orjson',
    'casbin', 'sqlalchemy_adapter'
}


class _MockLoader(importlib.abc.Loader):

    def create_module(self, spec):
        module = mock.MagicMock(name=spec.name)
        module.__name__ = spec.name
        module.__path__ = []
        module.__spec__ = spec
        module.__loader__ = self
        return module

    def exec_module(self, module):
        del module


class _MockFinder(importlib.abc.MetaPathFinder):
    """Resolves any import below a mocked top-level package to a MagicMock."""

    def find_spec(self, fullname, path, target=None):
        del path, target
        if fullname.split('.')[0] in _MOCKED:
            return importlib.machinery.ModuleSpec(fullname,
                                                  _MockLoader(),
                                                  is_package=True)
        return None


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'sky'))


def write_catalogs(home: str, catalogs: Dict[str, 'object'],
                   accelerators: Optional[Dict[str, Sequence[str]]] = None):
    """Places `{cloud: DataFrame}` as `vms.csv` files below a scratch $HOME."""
    root = os.path.join(home, '.sky', 'catalogs', 'v8')
    for cloud, df in catalogs.items():
        os.makedirs(os.path.join(root, cloud), exist_ok=True)
        df.to_csv(os.path.join(root, cloud, 'vms.csv'), index=False)
    os.makedirs(os.path.join(root, 'common'), exist_ok=True)
    # accelerator_registry reads AcceleratorName,Clouds
    # (reference sky/utils/accelerator_registry.py:38, :96-102).
    lines = ['AcceleratorName,Clouds']
    if accelerators is None:
        names = {}
        for cloud, df in catalogs.items():
            for name in df['AcceleratorName'].dropna().unique():
                names.setdefault(str(name), []).append(cloud)
        accelerators = names
    for name, clouds in sorted(accelerators.items()):
        lines.append(f'{name},"{list(clouds)!r}"')
    with open(os.path.join(root, 'common', 'accelerators.csv'),
              'w',
              encoding='utf-8') as f:
        f.write('\n'.join(lines) + '\n')
    # device memory table behind '32GB+' requests
    # (sky/utils/accelerator_registry.py:39, :50-73)
    from skypilot_b200 import synth  # pylint: disable=import-outside-toplevel
    synth.accelerator_metadata().to_csv(
        os.path.join(root, 'common', 'metadata.csv'), index=False)
    # <cloud>/images.csv behind `skypilot:` image tags
    # (sky/catalog/aws_catalog.py:87, :355-372)
    for cloud, frame in synth.images(catalogs).items():
        frame.to_csv(os.path.join(root, cloud, 'images.csv'), index=False)


def import_reference(home: str, enabled_clouds: Sequence[str]):
    """Imports `sky` from the reference tree with `$HOME=home`.

    Must be called once per process, before any `sky` import. Returns the
    imported `sky` module with the enabled-cloud cache patched.
    """
    if not reference_available():
        raise RuntimeError(f'reference tree not found at {REFERENCE_ROOT}')
    os.environ['HOME'] = home
    os.environ['PYTHONDONTWRITEBYTECODE'] = '1'
    os.environ['SKYPILOT_DISABLE_USAGE_COLLECTION'] = '1'
    os.environ['SKYPILOT_SKIP_CLOUD_IDENTITY_CHECK'] = '1'
    sys.dont_write_bytecode = True
    for p in (REFERENCE_ROOT, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if not any(isinstance(f, _MockFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _MockFinder())
    import sky  # pylint: disable=import-outside-toplevel
    from sky import check as sky_check  # pylint: disable=import-outside-toplevel
    from sky import clouds as sky_clouds  # pylint: disable=import-outside-toplevel
    from sky.utils import registry  # pylint: disable=import-outside-toplevel

    cloud_objs = [
        registry.CLOUD_REGISTRY.from_str(name) for name in enabled_clouds
    ]
    assert all(c is not None for c in cloud_objs), enabled_clouds
    del sky_clouds

    def _enabled(*args, **kwargs):
        del args, kwargs
        return list(cloud_objs)

    sky_check.get_cached_enabled_clouds_or_refresh = _enabled
    sky_check.check_capability = lambda *a, **k: None
    return sky


def clear_request_cache():
    from sky.utils import annotations  # pylint: disable=import-outside-toplevel
    annotations.clear_request_level_cache()


def as_namespace(**kwargs) -> types.SimpleNamespace:
    return types.SimpleNamespace(**kwargs)
