"""GPU-backed implementations of the catalog queries.

Same function names and return conventions as the reference's pandas
implementations in sky/catalog/common.py; the first argument is a
`CatalogView` (a cloud of the active store plus the row restrictions the
per-cloud wrapper applied) or, for the reference's own inline-DataFrame tests,
a pandas DataFrame that is ingested on the fly.

Row-level work runs on the device (`engine.scan` / `engine.solve`); what stays
on the host are dictionary look-ups keyed by instance-type / region name.
"""
import math
import threading
import weakref
from typing import Dict, List, NamedTuple, Optional, Tuple, Union

import numpy as np
import pandas as pd

from skypilot_b200 import _native
from skypilot_b200 import engine
from skypilot_b200.catalog.store import CatalogStore
from skypilot_b200.clouds import cloud as cloud_lib

_INLINE_CLOUD = 'inline'
_inline_lock = threading.Lock()
_inline_stores: Dict[int, Tuple['weakref.ref', CatalogStore]] = {}


class InstanceTypeInfo(NamedTuple):
    """Instance type information (sky/catalog/common.py:38-62)."""
    cloud: str
    instance_type: Optional[str]
    accelerator_name: str
    accelerator_count: float
    cpu_count: Optional[float]
    device_memory: Optional[float]
    memory: Optional[float]
    price: float
    spot_price: float
    region: str


class CatalogView:
    """A cloud's rows with the wrapper-level restrictions applied lazily.

    `flags_require` / `group` play the role of the reference's
    `df[df['InstanceType'].str.startswith(prefixes)]` pre-filters; they are
    evaluated by the scan kernel from the per-row flag column.
    """

    def __init__(self, store: CatalogStore, cloud: str,
                 flags_require: int = 0, group: int = 0,
                 local_disk: Optional[str] = None, device: int = 0):
        self.store = store
        self.cloud = cloud
        self.flags_require = flags_require
        self.group = group
        self.local_disk = local_disk
        self.device = device

    def restrict(self, flags_require: int = 0, group: int = 0,
                 local_disk: Optional[str] = None) -> 'CatalogView':
        return CatalogView(self.store, self.cloud,
                           self.flags_require | flags_require,
                           group or self.group, local_disk or self.local_disk,
                           self.device)

    @property
    def table(self):
        return self.store.cloud(self.cloud)


def _as_view(df_or_view: Union[CatalogView, pd.DataFrame]) -> CatalogView:
    if isinstance(df_or_view, CatalogView):
        return df_or_view
    df = df_or_view
    with _inline_lock:
        hit = _inline_stores.get(id(df))
        if hit is not None and hit[0]() is df:
            return CatalogView(hit[1], _INLINE_CLOUD)
        store = CatalogStore.from_frames({_INLINE_CLOUD: df})
        key = id(df)

        def _drop(_ref, key=key):
            with _inline_lock:
                _inline_stores.pop(key, None)

        _inline_stores[key] = (weakref.ref(df, _drop), store)
    return CatalogView(store, _INLINE_CLOUD)


def filter_with_local_disk(df_or_view, local_disk: Optional[str]):
    """Restricts a view to rows with the requested local disk
    (sky/catalog/common.py:481-506); evaluated inside the scan."""
    view = _as_view(df_or_view)
    if local_disk is None:
        return view
    return view.restrict(local_disk=local_disk)


def get_instance_type_for_cpus_mem_impl(
        df_or_view,
        cpus: Optional[str],
        memory_gb_or_ratio: Optional[str],
        region: Optional[str] = None,
        zone: Optional[str] = None,
        use_spot: bool = False,
        max_hourly_cost: Optional[float] = None) -> Optional[str]:
    """Cheapest instance type meeting the vCPU / memory request
    (sky/catalog/common.py:518-569)."""
    view = _as_view(df_or_view)
    b = engine.ProblemBuilder(view.store)
    b.add_query(
        b.cpus_mem_query(view.cloud, cpus, memory_gb_or_ratio, region, zone,
                         use_spot, max_hourly_cost,
                         flags_require=view.flags_require, group=view.group,
                         local_disk=view.local_disk))
    out = engine.scan(b, device=view.device)
    inst = int(out.results['best_inst'][0])
    if inst < 0:
        return None
    return view.store.inst_names[inst]


def get_instance_type_for_accelerator_impl(
    df_or_view,
    acc_name: str,
    acc_count: Union[int, float],
    cpus: Optional[str] = None,
    memory: Optional[str] = None,
    use_spot: bool = False,
    region: Optional[str] = None,
    zone: Optional[str] = None,
    max_hourly_cost: Optional[float] = None,
    flags_require2: int = 0,
) -> Tuple[Optional[List[str]], List[str]]:
    """Instance types with the accelerator, sorted by price, plus fuzzy
    candidates when nothing matches (sky/catalog/common.py:641-694)."""
    view = _as_view(df_or_view)
    b = engine.ProblemBuilder(view.store)
    spec = b.accelerator_query(view.cloud, acc_name, acc_count, cpus, memory,
                               use_spot, region, zone, max_hourly_cost,
                               local_disk=view.local_disk,
                               flags_require2=flags_require2,
                               want_list=True, want_fuzzy=True)
    spec['flags_require'] |= view.flags_require
    b.add_query(spec)
    table = view.table
    n_inst = max(len(table.inst_names), 1)
    n_keys = max(len(view.store.acc_keys), 1)
    out = engine.scan(b, list_cap=min(n_inst, 2048),
                      fuzzy_cap=min(n_keys, 2048), device=view.device)
    res = out.results[0]
    if not res['any_stage1']:
        return None, engine.format_fuzzy(view.store, out.fuzzy_list(0))
    if res['best_inst'] < 0:
        # rows matched, but none survives cpus / memory / price
        return [], []
    return [view.store.inst_names[i] for i in out.instance_list(0)], []


def _expand(view: CatalogView, instance_type: str, use_spot: bool,
            region: Optional[str], zone: Optional[str], split_by_zone: bool,
            us_first: bool = False):
    """Runs the expansion kernel for one instance type; returns the ordered
    candidate table (region id, zone id, hourly price)."""
    store = view.store
    table = view.table
    inst = table.inst_index.get(instance_type, -1)
    if inst < 0:
        return None
    b = engine.ProblemBuilder(store)
    s = b.add_slot(cloud=table.index, inst_id=inst,
                   price_col=1 if use_spot else 0,
                   region_id=engine.region_filter_id(table, region),
                   zone_id=engine.zone_exact_id(table, zone),
                   split_by_zone=int(split_by_zone), us_first=int(us_first),
                   use_spot=int(use_spot))
    t = b.add_task(s, s + 1)
    b.add_dag(t, t + 1, True, True)
    sol = engine.solve(b, device=view.device, want_tables=True)
    return sol.task_table(0)


def get_hourly_cost_impl(df_or_view, instance_type: str, use_spot: bool,
                         region: Optional[str], zone: Optional[str]) -> float:
    """Hourly price of an instance type in a region / zone
    (sky/catalog/common.py:360-400)."""
    view = _as_view(df_or_view)
    cands = _expand(view, instance_type, use_spot, region, zone,
                    split_by_zone=zone is not None)
    if cands is None or len(cands) == 0:
        table = view.table
        inst = table.inst_index.get(instance_type, -1)
        exists = inst >= 0 and _has_rows(view, inst, region, zone)
        if not exists:
            if zone is None:
                where = ('all regions'
                         if region is None else f'region {region!r}')
            else:
                where = f'zone {zone!r}'
            raise ValueError(
                f'Instance type {instance_type!r} not found in {where}.')
        price_str = 'SpotPrice' if use_spot else 'Price'
        raise ValueError(
            f'No {price_str} found for instance type {instance_type!r}.')
    return float(np.min(cands['hourly']))


def _has_rows(view: CatalogView, inst: int, region: Optional[str],
              zone: Optional[str]) -> bool:
    """Error-path only: does (instance type, region, zone) have any row?"""
    c = view.store.columns
    table = view.table
    rows = c['inst_rows'][c['inst_row_offsets'][inst]:
                          c['inst_row_offsets'][inst + 1]]
    ok = np.ones(len(rows), dtype=bool)
    if region is not None:
        ok &= c['region_id'][rows] == engine.region_filter_id(table, region)
    if zone is not None:
        ok &= c['zone_id'][rows] == engine.zone_exact_id(table, zone)
    return bool(ok.any())


def get_region_zones_for_instance_type_impl(
        df_or_view, instance_type: str, use_spot: bool,
        us_first: bool = False) -> List[cloud_lib.Region]:
    """Regions (with zones) offering an instance type, cheapest first
    (sky/catalog/common.py:793-809 + aws_catalog.py:322-336)."""
    view = _as_view(df_or_view)
    cands = _expand(view, instance_type, use_spot, None, None,
                    split_by_zone=True, us_first=us_first)
    return regions_from_candidates(view.table, cands)


def regions_from_candidates(table, cands) -> List[cloud_lib.Region]:
    regions: List[cloud_lib.Region] = []
    if cands is None:
        return regions
    by_id: Dict[int, cloud_lib.Region] = {}
    for rid, zid in zip(cands['region_id'], cands['zone_id']):
        rid = int(rid)
        region = by_id.get(rid)
        if region is None:
            region = cloud_lib.Region(table.region_names[rid])
            if table.has_zone_column:
                region.set_zones([])
            by_id[rid] = region
            regions.append(region)
        if table.has_zone_column and zid >= 0:
            zone = cloud_lib.Zone(table.zone_names[int(zid)])
            zone.region = region
            region.zones.append(zone)
    return regions


# ---- dictionary look-ups (host metadata, no row scans) ----------------------
def instance_type_exists_impl(df_or_view, instance_type: str) -> bool:
    return instance_type in _as_view(df_or_view).table.inst_index


def _value(v) -> Optional[float]:
    if v is None or (isinstance(v, float) and math.isnan(v)) or pd.isna(v):
        return None
    return float(v)


def get_vcpus_mem_from_instance_type_impl(
        df_or_view,
        instance_type: str) -> Tuple[Optional[float], Optional[float]]:
    view = _as_view(df_or_view)
    try:
        row = view.store.instance_row(view.cloud, instance_type)
    except KeyError:
        raise ValueError(f'No instance type {instance_type} found.') from None
    return _value(row.get('vCPUs')), _value(row.get('MemoryGiB'))


def get_accelerators_from_instance_type_impl(
        df_or_view,
        instance_type: str) -> Optional[Dict[str, Union[int, float]]]:
    view = _as_view(df_or_view)
    try:
        row = view.store.instance_row(view.cloud, instance_type)
    except KeyError:
        raise ValueError(f'No instance type {instance_type} found.') from None
    name, count = row.get('AcceleratorName'), row.get('AcceleratorCount')
    if name is None or pd.isnull(name):
        return None
    count = float(count)
    return {name: int(count) if int(count) == count else count}


def get_arch_from_instance_type_impl(df_or_view,
                                     instance_type: str) -> Optional[str]:
    view = _as_view(df_or_view)
    try:
        row = view.store.instance_row(view.cloud, instance_type)
    except KeyError:
        raise ValueError(f'No instance type {instance_type} found.') from None
    arch = row.get('Arch')
    if arch is None or pd.isnull(arch):
        return None
    return arch


def get_local_disk_from_instance_type_impl(
        df_or_view, instance_type: str) -> Optional[str]:
    view = _as_view(df_or_view)
    try:
        row = view.store.instance_row(view.cloud, instance_type)
    except KeyError:
        raise ValueError(f'No instance type {instance_type} found.') from None
    mode = row.get('LocalDiskType')
    if mode is None or pd.isna(mode) or mode != 'ssd':
        return None
    size, count = row.get('LocalDiskSize'), row.get('LocalDiskCount')
    if size is None or count is None or pd.isna(size) or pd.isna(count):
        return None
    nvme = row.get('NVMeSupported', False)
    if nvme is not None and not pd.isna(nvme) and bool(nvme):
        mode = 'nvme'
    return f'{mode}:{int(float(size) * float(count))}'


def validate_region_zone_impl(
        cloud_name: str, df_or_view, region: Optional[str],
        zone: Optional[str]) -> Tuple[Optional[str], Optional[str]]:
    """Canonical (region, zone) or ValueError
    (sky/catalog/common.py:289-357)."""
    import difflib  # pylint: disable=import-outside-toplevel
    table = _as_view(df_or_view).table

    def _candidates(loc: str, all_loc: List[str]) -> str:
        close = sorted(difflib.get_close_matches(loc, all_loc, n=5,
                                                 cutoff=0.9))
        if not close:
            return ''
        return f'\nDid you mean one of these: {", ".join(close)!r}?'

    valid_region, valid_zone = region, zone
    region_id = None
    if region is not None:
        region_id = table.region_lower.get(region.lower())
        if region_id is None:
            msg = f'Invalid region {region!r}'
            hint = _candidates(region.lower(), list(table.region_lower))
            if not hint:
                regions = ', '.join(sorted(table.region_lower))
                if cloud_name in ('azure', 'gcp'):
                    msg += (
                        '\nIf a region is not included in the following '
                        'list, please check the FAQ docs for how to fetch '
                        'its catalog info.\nhttps://docs.skypilot.co'
                        '/en/latest/reference/faq.html#advanced-how-to-'
                        'make-skypilot-use-all-global-regions')
                msg += (f'\nList of supported {cloud_name} regions: '
                        f'{regions!r}')
            raise ValueError(msg + hint)
        valid_region = table.region_names[region_id]
    if zone is not None:
        zone_id = table.zone_exact.get(zone)
        if zone_id is not None and region_id is not None and (
                table.zone_region[zone_id] != region_id):
            zone_id = None
        if zone_id is None:
            region_str = f' for region {region!r}' if region else ''
            pool = [
                z for i, z in enumerate(table.zone_names)
                if region_id is None or table.zone_region[i] == region_id
            ]
            raise ValueError(f'Invalid zone {zone!r}{region_str}' +
                             _candidates(zone, pool))
        valid_region = table.region_names[table.zone_region[zone_id]]
    return valid_region, valid_zone


def get_image_id_from_tag_impl(df: pd.DataFrame, tag: str,
                               region: Optional[str]) -> Optional[str]:
    """The image id of `tag` in `region` (sky/catalog/common.py:813-831):
    with region None the tag must name a single image; None when nothing
    matches or the id is missing."""
    df = df[df['Tag'] == tag]
    if region is not None:
        df = df[df['Region'].str.lower() == region.lower()]
    assert len(df) <= 1, ('Multiple images found for tag '
                          f'{tag} in region {region}')
    if df.empty:
        return None
    image_id = df['ImageId'].iloc[0]
    if pd.isna(image_id):
        return None
    return image_id


def is_image_tag_valid_impl(df: pd.DataFrame, tag: str,
                            region: Optional[str]) -> bool:
    """sky/catalog/common.py:834-840."""
    df = df[df['Tag'] == tag]
    if region is not None:
        df = df[df['Region'].str.lower() == region.lower()]
    df = df.dropna(subset=['ImageId'])
    return not df.empty
