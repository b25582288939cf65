"""GCP catalog: the function table of sky/catalog/gcp_catalog.py, served from
the GPU-resident store.

GCP keeps accelerators in their own catalog rows (no InstanceType) and bills
them on top of a host VM, so three functions differ from the generic table:
the host-VM choice of `get_instance_type_for_accelerator`
(gcp_catalog.py:334-393), `get_accelerator_hourly_cost` (:424-442) and
`get_region_zones_for_accelerators` (:574-581).
"""
from typing import Dict, List, Optional, Tuple

import numpy as np

from skypilot_b200 import _native
from skypilot_b200 import engine
from skypilot_b200.catalog import _cloud_catalog
from skypilot_b200.catalog import common
from skypilot_b200.catalog import rules

_impl = _cloud_catalog.CloudCatalog('gcp', supports_zones=True)

validate_region_zone = _impl.validate_region_zone
get_default_instance_type = _impl.get_default_instance_type
get_arch_from_instance_type = _impl.get_arch_from_instance_type
get_local_disk_from_instance_type = _impl.get_local_disk_from_instance_type
get_region_zones_for_instance_type = _impl.get_region_zones_for_instance_type

GCP_ACC_INSTANCE_TYPES = list(rules.GCP_INSTANCE_TO_ACC.keys())


def instance_type_exists(instance_type: str) -> bool:
    if instance_type == 'TPU-VM':
        return True
    return _impl.instance_type_exists(instance_type)


def get_hourly_cost(instance_type: str, use_spot: bool = False,
                    region: Optional[str] = None,
                    zone: Optional[str] = None) -> float:
    if instance_type == 'TPU-VM':
        return 0  # the TPU VM host is not billed separately
    return _impl.get_hourly_cost(instance_type, use_spot, region, zone)


def get_vcpus_mem_from_instance_type(
        instance_type: str) -> Tuple[Optional[float], Optional[float]]:
    if instance_type == 'TPU-VM':
        return None, None
    return _impl.get_vcpus_mem_from_instance_type(instance_type)


def get_accelerators_from_instance_type(
        instance_type: str) -> Optional[Dict[str, int]]:
    return rules.GCP_INSTANCE_TO_ACC.get(instance_type)


def get_instance_type_for_accelerator(
        acc_name: str, acc_count: int, cpus: Optional[str] = None,
        memory: Optional[str] = None, use_spot: bool = False,
        local_disk: Optional[str] = None, region: Optional[str] = None,
        zone: Optional[str] = None, max_hourly_cost: Optional[float] = None
) -> Tuple[Optional[List[str]], List[str]]:
    """([host VM type], []) or (None, fuzzy candidates)."""
    del local_disk
    view = _impl._view()  # pylint: disable=protected-access
    b = engine.ProblemBuilder(view.store)
    gate = b.add_query(
        b.accelerator_query('gcp', acc_name, acc_count, cpus, memory, use_spot,
                            region, zone, max_hourly_cost, want_list=False,
                            want_fuzzy=True))
    host = None
    if acc_name in rules.GCP_FIXED_HOSTS:
        group = rules.GCP_GROUP_IDS.get((acc_name, acc_count))
        if group is not None:
            host = b.add_query(
                b.cpus_mem_query('gcp', cpus, memory, group=group))
    else:
        table = rules.GCP_ACC_HOST_CPUS.get(acc_name,
                                            rules.GCP_ACC_HOST_CPUS['DEFAULT'])
        default_cpus = table.get(acc_count)
        no_rule = cpus is None and memory is None and default_cpus is None
        if not no_rule:
            if cpus is None and memory is None:
                cpus = f'{default_cpus}+'
            if memory is None:
                assert cpus is not None, (acc_name, acc_count)
                cpu_val = int(cpus.strip('+').strip('x'))
                memory = f'{cpu_val * rules.GCP_GPU_MEMORY_CPU_RATIO}+'
            host = b.add_query(
                b.cpus_mem_query('gcp', cpus, memory,
                                 flags_require=_native.F_HOST_FAMILY))
    out = engine.scan(b, fuzzy_cap=min(max(len(view.store.acc_keys), 1),
                                       2048), device=view.device)
    if not out.results['any_stage1'][gate]:
        return None, engine.format_fuzzy(view.store, out.fuzzy_list(gate))
    if acc_name not in rules.GCP_FIXED_HOSTS:
        # the reference only looks the host rule up once accelerator rows
        # exist, and asserts there (gcp_catalog.py:376-378)
        assert host is not None, (acc_name, acc_count)
    if host is None or out.results['best_inst'][host] < 0:
        return None, []
    return [view.store.inst_names[int(out.results['best_inst'][host])]], []


def _accelerator_table(accelerator: str, count, use_spot: bool,
                       region: Optional[str], zone: Optional[str],
                       instance_type: Optional[str]):
    view = _impl._view()  # pylint: disable=protected-access
    table = view.table
    _, _, strict = engine.accelerator_sets(view.store, accelerator, count)
    b = engine.ProblemBuilder(view.store)
    inst = -2
    if instance_type is not None and instance_type != 'TPU-VM':
        inst = table.inst_index.get(instance_type, -1)
        if inst < 0:
            return view, None
    s = b.add_slot(cloud=table.index, inst_id=inst, acc_words=strict,
                   price_col=1 if use_spot else 0,
                   region_id=engine.region_exact_id(table, region),
                   zone_id=engine.zone_exact_id(table, zone),
                   split_by_zone=1, use_spot=int(use_spot))
    t = b.add_task(s, s + 1)
    b.add_dag(t, t + 1, True, True)
    sol = engine.solve(b, device=view.device, want_tables=True)
    return view, sol.task_table(0)


def get_region_zones_for_accelerators(accelerator: str, count: int,
                                      use_spot: bool = False,
                                      instance_type: Optional[str] = None):
    """Regions / zones offering the accelerator, cheapest first; with
    `instance_type`, only zones that also offer that host VM -- the
    intersection of sky/clouds/gcp.py:296-322, done by the expansion kernel."""
    view, cands = _accelerator_table(accelerator, count, use_spot, None, None,
                                     instance_type)
    return common.regions_from_candidates(view.table, cands)


def get_accelerator_hourly_cost(accelerator: str, count: int,
                                use_spot: bool = False,
                                region: Optional[str] = None,
                                zone: Optional[str] = None) -> float:
    """Hourly price of the accelerators alone (gcp_catalog.py:424-442): the
    spot price falls back to on-demand when no zone lists one."""
    _, cands = _accelerator_table(accelerator, count, use_spot, region, zone,
                                  None)
    if use_spot and (cands is None or len(cands) == 0):
        _, cands = _accelerator_table(accelerator, count, False, region, zone,
                                      None)
    if cands is None or len(cands) == 0:
        # the reference asserts one price per region on the (empty) frame
        assert region is None, (
            f'no {accelerator}:{count} rows in {region} / {zone}')
        return float('nan')
    # `hourly` = host (0 here) + accelerator price
    return float(np.min(cands['hourly']))


def list_accelerators(gpus_only: bool, name_filter: Optional[str] = None,
                      region_filter: Optional[str] = None,
                      quantity_filter: Optional[int] = None,
                      case_sensitive: bool = True, all_regions: bool = False,
                      require_price: bool = True):
    """GPUs / TPUs offered by GCP with the price of accelerator + cheapest
    host VM of the zone (gcp_catalog.py:445-571)."""
    from skypilot_b200.catalog import listing  # pylint: disable=import-outside-toplevel
    return listing.gcp_listing(_impl._view(), gpus_only, name_filter,  # pylint: disable=protected-access
                               region_filter, quantity_filter, case_sensitive,
                               all_regions, require_price)



def check_accelerator_attachable_to_host(instance_type: str,
                                         accelerators: Optional[Dict[str, int]],
                                         zone: Optional[str] = None) -> None:
    """Can the accelerators be attached to this host VM? The documented
    limits of GCP (host families of A100 / L4 / H100 ..., valid counts,
    maximum vCPUs and memory of the N1 host per accelerator count):
    sky/catalog/gcp_catalog.py:584-683, same checks in the same order, same
    messages. Raises exceptions.ResourcesMismatchError."""
    from skypilot_b200 import exceptions  # pylint: disable=import-outside-toplevel
    if accelerators is None:
        if instance_type in rules.GCP_INSTANCE_TO_ACC:
            accelerators = rules.GCP_INSTANCE_TO_ACC[instance_type]
        else:
            return
    acc = list(accelerators.items())
    assert len(acc) == 1, acc
    acc_name, acc_count = acc[0]
    if not list_accelerators(gpus_only=False, name_filter=acc_name):
        raise exceptions.ResourcesMismatchError(
            f'{acc_name} is not available in GCP. '
            'See \'sky gpus list --cloud gcp\'')
    if acc_name.startswith('tpu-'):
        if instance_type != 'TPU-VM' and not instance_type.startswith('n1-'):
            raise exceptions.ResourcesMismatchError(
                'TPU Nodes can be only used with N1 machines. '
                'Please refer to: '
                'https://cloud.google.com/compute/docs/general-purpose-machines#n1_machines')  # pylint: disable=line-too-long
        return
    fixed = rules.GCP_FIXED_HOSTS
    limits = rules.GCP_ACC_MAX_CPU_MEM
    if acc_name in fixed:
        matching_types = fixed[acc_name].get(acc_count)
        if matching_types is None:
            raise KeyError(acc_count)  # as the reference's dict look-up does
        if instance_type not in matching_types:
            raise exceptions.ResourcesMismatchError(
                f'{acc_name} GPUs cannot be attached to {instance_type}. '
                f'Use one of {matching_types} instead. Please refer to '
                'https://cloud.google.com/compute/docs/gpus')
    elif not instance_type.startswith('n1-'):
        raise exceptions.ResourcesMismatchError(
            f'{acc_name} GPUs cannot be attached to {instance_type}. '
            'Use N1 instance types instead. Please refer to: '
            'https://cloud.google.com/compute/docs/machine-types#gpus')
    if acc_name in fixed:
        valid_counts = list(fixed[acc_name].keys())
    else:
        assert acc_name in limits, acc_name
        valid_counts = list(limits[acc_name].keys())
    if acc_count not in valid_counts:
        raise exceptions.ResourcesMismatchError(
            f'{acc_name}:{acc_count} is not launchable on GCP. '
            f'The valid {acc_name} counts are {valid_counts}.')
    if acc_name in fixed:
        max_cpus, max_memory = get_vcpus_mem_from_instance_type(instance_type)
    else:
        max_cpus, max_memory = limits[acc_name][acc_count]
        if acc_name == 'K80' and acc_count == 8:
            if zone in ['asia-east1-a', 'us-east1-d']:
                max_memory = 416
        elif acc_name == 'P100' and acc_count == 4:
            if zone in ['us-east1-c', 'europe-west1-d', 'europe-west1-b']:
                max_cpus = 64
                max_memory = 208
    num_cpus, memory = get_vcpus_mem_from_instance_type(instance_type)
    if num_cpus > max_cpus:
        raise exceptions.ResourcesMismatchError(
            f'{acc_name}:{acc_count} cannot be attached to '
            f'{instance_type}. The maximum number of vCPUs is {max_cpus}. '
            'Please refer to: https://cloud.google.com/compute/docs/gpus')
    if memory > max_memory:
        raise exceptions.ResourcesMismatchError(
            f'{acc_name}:{acc_count} cannot be attached to '
            f'{instance_type}. The maximum CPU memory is {max_memory} GB. '
            'Please refer to: https://cloud.google.com/compute/docs/gpus')


def get_image_id_from_tag(tag: str, region: Optional[str] = None
                         ) -> Optional[str]:
    return _impl.get_image_id_from_tag(tag, region)


def is_image_tag_valid(tag: str, region: Optional[str]) -> bool:
    """GCP images are not region-specific (gcp_catalog.py:699-704)."""
    del region
    return get_image_id_from_tag(tag, None) is not None
