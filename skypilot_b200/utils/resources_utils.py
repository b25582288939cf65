"""Small value types shared by Resources / clouds / optimizer.

Mirrors the parts of sky/utils/resources_utils.py the placement path touches:
DiskTier / NetworkTier (:30-75), local-disk strings (:111-239),
FeasibleResources (:407-422), memory strings (:505-570) and
make_launchables_for_valid_region_zones (:454-502).
"""
import dataclasses
import enum
from typing import List, Optional, Tuple, Union

_UNITS = {
    'kb': 2**10, 'ki': 2**10, 'mb': 2**20, 'mi': 2**20, 'gb': 2**30,
    'gi': 2**30, 'tb': 2**40, 'ti': 2**40, 'pb': 2**50, 'pi': 2**50,
}
DEFAULT_LOCAL_DISK_SIZE = '100+'


class DiskTier(enum.Enum):
    LOW = 'low'
    MEDIUM = 'medium'
    HIGH = 'high'
    ULTRA = 'ultra'
    BEST = 'best'

    def __le__(self, other: 'DiskTier') -> bool:
        order = list(DiskTier)
        return order.index(self) <= order.index(other)


class NetworkTier(enum.Enum):
    STANDARD = 'standard'
    BEST = 'best'


@dataclasses.dataclass
class FeasibleResources:
    """What a cloud answers to "can you run this request?"."""
    resources_list: List['object']
    fuzzy_candidate_list: List[str]
    hint: Optional[str]


def parse_memory_resource(value: Union[str, int, float],
                          field_name: str,
                          ret_type: type = int,
                          unit: str = 'gb',
                          allow_plus: bool = False,
                          allow_x: bool = False) -> str:
    """'16', '16+', '4x', '32GB', '2048mb+' -> the quantity in `unit`."""
    text = str(value)
    error = (f'"{field_name}" field should be a <number>[unit][+], '
             f'got {value}')
    plus = x = ''
    if text.endswith('+'):
        if not allow_plus:
            raise ValueError(error)
        text, plus = text[:-1], '+'
    if text.endswith('x'):
        if not allow_x:
            raise ValueError(error)
        text, x = text[:-1], 'x'
    try:
        ret_type(text)
        return f'{text}{plus}{x}'
    except ValueError:
        pass
    low = text.lower()
    for suffix, mult in _UNITS.items():
        if low.endswith(suffix):
            try:
                number = ret_type(low[:-len(suffix)])
            except ValueError:
                continue
            converted = number * mult / _UNITS[unit]
            if ret_type(converted) != converted:
                raise ValueError(error)
            return f'{ret_type(converted)}{plus}{x}'
    raise ValueError(error)


def normalize_local_disk(local_disk: str) -> str:
    """'nvme', '1000+', 'ssd:500' -> 'mode:size[+]'."""
    text = str(local_disk).lower().strip()
    parts = text.split(':')

    def check(size: str) -> None:
        try:
            if float(size.rstrip('+')) <= 0:
                raise ValueError
        except ValueError:
            raise ValueError(
                f'Invalid local_disk: {text!r}. Expected "mode:size[+]", '
                '"mode", or "size[+]". Mode must be "nvme" or "ssd", size '
                'must be positive (GB), optionally with "+".') from None

    if len(parts) == 1:
        if parts[0] in ('nvme', 'ssd'):
            return f'{parts[0]}:{DEFAULT_LOCAL_DISK_SIZE}'
        check(parts[0])
        return f'nvme:{parts[0]}'
    if len(parts) == 2:
        if parts[0] not in ('nvme', 'ssd'):
            raise ValueError(f'Invalid local_disk mode: {parts[0]!r}. '
                             'Must be "nvme" or "ssd".')
        check(parts[1])
        return f'{parts[0]}:{parts[1]}'
    raise ValueError(f'Invalid local_disk format: {text!r}.')


def parse_local_disk_str(local_disk: str) -> Tuple[str, float, bool]:
    mode, size = local_disk.split(':')[:2]
    at_least = size.endswith('+')
    return mode, float(size[:-1] if at_least else size), at_least


def make_launchables_for_valid_region_zones(
        launchable_resources,
        override_optimize_by_zone: bool = False) -> List['object']:
    """One launchable per region, or per zone for spot / zone-priced clouds."""
    assert launchable_resources.is_launchable()
    out = []
    for region in launchable_resources.get_valid_regions_for_launchable():
        by_zone = (override_optimize_by_zone or
                   launchable_resources.cloud.optimize_by_zone())
        if region.zones is not None and (launchable_resources.use_spot or
                                         by_zone):
            for zone in region.zones:
                out.append(
                    launchable_resources.copy(region=region.name,
                                              zone=zone.name))
        else:
            out.append(launchable_resources.copy(region=region.name))
    return out
