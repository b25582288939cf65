"""Accelerator metadata look-ups of sky/utils/accelerator_registry.py that
create request alternatives: which devices have a given memory size.

The reference reads `common/metadata.csv` (columns GPU, MemoryGB,
Manufacturer) of the catalog directory; here the table lives on the loaded
store (`CatalogStore.metadata`, read from `<catalog dir>/common/metadata.csv`
or set with `set_accelerator_metadata`)."""
from typing import List, Optional


def get_devices_by_memory(memory: float, plus: bool = False,
                          manufacturer: Optional[str] = None) -> List[str]:
    """Devices with MemoryGB == memory (>= with `plus`), optionally of one
    manufacturer, in table order (accelerator_registry.py:50-73)."""
    from skypilot_b200 import catalog  # pylint: disable=import-outside-toplevel
    store = catalog.get_store(required=False)
    df = getattr(store, 'metadata', None) if store is not None else None
    if df is None or df.empty:
        return []
    if plus:
        df = df[df['MemoryGB'] >= memory]
    else:
        df = df[df['MemoryGB'] == memory]
    if manufacturer is not None:
        df = df[df['Manufacturer'].str.lower() == manufacturer.lower()]
    return df['GPU'].tolist()
