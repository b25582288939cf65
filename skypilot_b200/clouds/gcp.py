"""Google Cloud Platform (placement-relevant part of sky/clouds/gcp.py).

GCP differs from the single-table clouds: accelerators are separate catalog
rows (no InstanceType) billed on top of a host VM, candidates are always per
zone, and the launchable zones are the accelerator's zones that also offer
the host VM (gcp.py:281-331, :709-823; gcp_catalog.py:334-442).
"""
from typing import Any, Dict, Optional

from skypilot_b200 import _native
from skypilot_b200.catalog import rules
from skypilot_b200.clouds import cloud
from skypilot_b200.utils import registry


def is_tpu(resources: Optional[Any]) -> bool:
    if resources is None or resources.accelerators is None:
        return False
    acc = list(resources.accelerators.keys())[0]
    return acc.startswith('tpu')


def is_tpu_vm(resources: Optional[Any]) -> bool:
    """TPU VM unless `accelerator_args: {tpu_vm: False}`
    (sky/clouds/utils/gcp_utils.py:37-47)."""
    if not is_tpu(resources):
        return False
    args = resources.accelerator_args
    if args is None:
        return True
    return args.get('tpu_vm', True)


@registry.CLOUD_REGISTRY.register
class GCP(cloud.Cloud):
    _REPR = 'GCP'
    _CATALOG = 'gcp'
    _DEFAULT_IMAGE_GB = 50  # DEFAULT_GCP_IMAGE_GB, sky/clouds/gcp.py:96

    @classmethod
    def _unsupported_features_for_resources(cls, resources: Any,
                                            region: Optional[str] = None):
        del region
        unsupported = {}
        if is_tpu(resources) and not is_tpu_vm(resources):
            unsupported[cloud.CloudImplementationFeatures.MULTI_NODE] = (
                'TPU node does not support multi-node. Please set '
                'num_nodes to 1.')
        unsupported[cloud.CloudImplementationFeatures.LOCAL_DISK] = (
            'Local disk is not supported on GCP')
        return unsupported

    @classmethod
    def optimize_by_zone(cls) -> bool:
        return True

    def get_egress_cost(self, num_gigabytes: float) -> float:
        """Worldwide egress $/GB (gcp.py:395-404)."""
        if num_gigabytes <= 1024:
            return 0.12 * num_gigabytes
        if num_gigabytes <= 1024 * 10:
            return 0.11 * num_gigabytes
        return 0.08 * num_gigabytes

    def accelerators_to_hourly_cost(self, accelerators: Dict[str, int],
                                    use_spot: bool,
                                    region: Optional[str] = None,
                                    zone: Optional[str] = None) -> float:
        assert len(accelerators) == 1, accelerators
        acc, acc_count = list(accelerators.items())[0]
        return self._catalog_module().get_accelerator_hourly_cost(
            acc, acc_count, use_spot=use_spot, region=region, zone=zone)

    @classmethod
    def get_accelerators_from_instance_type(
            cls, instance_type: str) -> Optional[Dict[str, Any]]:
        # GCP attaches accelerators separately; only the fixed A2/G2/A3 hosts
        # imply one (gcp_catalog.py:313-331).
        return rules.GCP_INSTANCE_TO_ACC.get(instance_type)

    @classmethod
    def regions_with_offering(cls, instance_type, accelerators, use_spot,
                              region, zone, resources=None):
        del resources
        module = cls._catalog_module()
        if accelerators is None:
            regions = module.get_region_zones_for_instance_type(
                instance_type, use_spot)
        else:
            assert len(accelerators) == 1, accelerators
            acc, count = list(accelerators.items())[0]
            regions = module.get_region_zones_for_accelerators(
                acc, count, use_spot, instance_type=instance_type)
        if region is not None:
            regions = [r for r in regions if r.name == region]
        if zone is not None:
            for r in regions:
                assert r.zones is not None, r
                r.set_zones([z for z in r.zones if z.name == zone])
            regions = [r for r in regions if r.zones]
        return regions

    def plan_fast(self, store, resources: Any, num_nodes: int):
        """The two common shapes of plan_feasible below (vCPUs / memory; one
        non-TPU accelerator), same bytes, positional packs
        (cloud.Cloud.plan_fast)."""
        r = resources
        if (r._instance_type is not None or r._image_id is not None or  # pylint: disable=protected-access
                r._local_disk is not None or r._disk_tier is not None or  # pylint: disable=protected-access
                r._network_tier is not None or r._ports is not None or  # pylint: disable=protected-access
                r._accelerator_args is not None or  # pylint: disable=protected-access
                cloud._late('skypilot_config').has_config()):  # pylint: disable=protected-access
            return None
        if num_nodes > 1 or r._use_spot:  # pylint: disable=protected-access
            unsupported = self._unsupported_features_for_resources(r, None)
            feats = cloud.CloudImplementationFeatures
            if ((r._use_spot and feats.SPOT_INSTANCE in unsupported) or  # pylint: disable=protected-access
                    (num_nodes > 1 and feats.MULTI_NODE in unsupported)):
                return None
        engine = cloud._late('engine')  # pylint: disable=protected-access
        ctxs = store.__dict__.setdefault('_fast_ctx', {})
        ctx = ctxs.get(GCP)
        if ctx is None:
            gcp = GCP()
            ctx = {
                'table': self._view().table, 'rules': self._rules(),
                'make_cpu': lambda name, res: res.copy(
                    cloud=gcp, instance_type=name, accelerators=None,
                    cpus=None, memory=None),
                'gcp': gcp,
            }
            ctxs[GCP] = ctx
        table = ctx['table']
        index = table.index
        use_spot = bool(r._use_spot)  # pylint: disable=protected-access
        region, zone = r._region, r._zone  # pylint: disable=protected-access
        max_cost = r._max_hourly_cost  # pylint: disable=protected-access
        max_price = cloud._INF if max_cost is None else float(max_cost)  # pylint: disable=protected-access
        plan = cloud.SlotPlan()
        rec = cloud._Recorded()  # pylint: disable=protected-access
        qpack, spack = engine._QUERY_PACK, engine._SLOT_PACK  # pylint: disable=protected-access
        has_inst = _native.F_HAS_INSTANCE
        accelerators = r._accelerators  # pylint: disable=protected-access
        cpus, memory = r._cpus, r._memory  # pylint: disable=protected-access
        region_x = engine.region_exact_id(table, region)
        zone_x = engine.zone_exact_id(table, zone)
        if accelerators is None:
            if cpus is None and memory is None:
                cpus = f'{ctx["rules"].default_cpus}+'
            if memory is None:
                memory = f'{ctx["rules"].default_mem_ratio}x'
            cop, cval = engine.parse_cpus(cpus)
            mop, mval = engine.parse_memory(memory)
            rec.query_recs.append(qpack(
                index, 0, _native.F_DEFAULT_FAMILY | has_inst, 0, -1, -1,
                1 if (use_spot and max_cost is not None) else 0, cop, mop, 0,
                engine.region_filter_id(table, region),
                engine.zone_filter_id(table, zone), 0, cval, mval, 0,
                max_price))
            plan.list_query = 0
            plan.make = ctx['make_cpu']
            rec.slot_recs.append(spack(
                index, 0, -1, -1, -1, 1 if use_spot else 0, region_x, zone_x,
                1, 0, -1, int(use_spot), -1, 1.0, 1.0, 3600.0))
            plan.slot = 0
            return plan, rec
        assert len(accelerators) == 1, resources
        acc, acc_count = next(iter(accelerators.items()))
        if acc.startswith('tpu-'):
            return None
        exact, _, strict = engine.accelerator_sets(store, acc, acc_count)
        cop, cval = engine.parse_cpus(cpus)
        mop, mval = engine.parse_memory(memory)
        # 1) the gate: do accelerator rows exist at all
        rec.query_recs.append(qpack(
            index, _native.Q_ACC, 0, 0, rec.add_set(store, exact), -1,
            1 if use_spot else 0, cop, mop, 0,
            engine.region_filter_id(table, region),
            engine.zone_filter_id(table, zone), 0, cval, mval, 0, max_price))
        plan.gate_query = 0
        plan.fuzzy_query = 0
        gcp = ctx['gcp']
        acc_dict = {acc: acc_count}
        plan.make = lambda name, res: res.copy(
            cloud=gcp, instance_type=name, accelerators=acc_dict, cpus=None,
            memory=None)
        # 2) the host VM
        if acc in rules.GCP_FIXED_HOSTS:
            group = rules.GCP_GROUP_IDS.get((acc, acc_count))
            if group is None:
                return plan, rec
            rec.query_recs.append(qpack(
                index, 0, has_inst, group, -1, -1, 0, cop, mop, 0, -1, -1, 0,
                cval, mval, 0, cloud._INF))  # pylint: disable=protected-access
        else:
            table_cpus = rules.GCP_ACC_HOST_CPUS.get(
                acc, rules.GCP_ACC_HOST_CPUS['DEFAULT'])
            default_cpus = table_cpus.get(acc_count)
            if cpus is None and memory is None:
                if default_cpus is None:
                    return plan, rec
                cpus = f'{default_cpus}+'
            if memory is None:
                cpu_val = int(cpus.strip('+').strip('x'))
                memory = f'{cpu_val * rules.GCP_GPU_MEMORY_CPU_RATIO}+'
            hop, hval = engine.parse_cpus(cpus)
            hmop, hmval = engine.parse_memory(memory)
            rec.query_recs.append(qpack(
                index, 0, _native.F_HOST_FAMILY | has_inst, 0, -1, -1, 0, hop,
                hmop, 0, -1, -1, 0, hval, hmval, 0, cloud._INF))  # pylint: disable=protected-access
        plan.list_query = 1
        key = store.acc_key_index.get((acc, float(acc_count)), -2)
        rec.slot_recs.append(spack(
            index, 1, -1, 0, rec.add_set(store, strict),
            1 if use_spot else 0, region_x, zone_x, 1, 0, key, int(use_spot),
            -1, 1.0, 1.0, 3600.0))
        plan.slot = 0
        return plan, rec

    def plan_feasible(self, builder, resources: Any,
                      want_list: bool = False,
                      want_fuzzy=None) -> cloud.SlotPlan:
        if want_fuzzy is None:
            want_fuzzy = want_list
        engine = cloud._late('engine')  # pylint: disable=protected-access
        view = self._view()
        table = view.table
        store = view.store
        plan = cloud.SlotPlan()
        use_spot = bool(resources.use_spot)
        slot_common = dict(
            cloud=table.index, price_col=1 if use_spot else 0,
            region_id=engine.region_exact_id(table, resources.region),
            zone_id=engine.zone_exact_id(table, resources.zone),
            split_by_zone=1, us_first=0, use_spot=int(use_spot),
            region_words=cloud.region_allow_words(table, resources, self))

        def acc_slot_fields(acc: str, count) -> Dict[str, Any]:
            _, _, strict = engine.accelerator_sets(store, acc, count)
            key = store.acc_key_index.get((acc, float(count)), -2)
            return dict(acc_words=strict, cand_acc_key=key)

        if resources.instance_type is not None:
            plan.explicit_instance = resources.instance_type
            plan.make = lambda name, res: res
            fields = dict(slot_common)
            if resources._accelerators is not None:  # pylint: disable=protected-access
                acc, count = list(resources.accelerators.items())[0]
                fields.update(acc_slot_fields(acc, count))
            if resources.instance_type == 'TPU-VM':
                fields['inst_id'] = -2
            else:
                inst = table.inst_index.get(resources.instance_type, -1)
                if inst < 0:
                    return plan
                fields['inst_id'] = inst
            plan.slot = builder.add_slot(**fields)
            return plan

        gcp = GCP()
        if resources.accelerators is None:
            cpus, memory = resources.cpus, resources.memory
            if cpus is None and memory is None:
                cpus = f'{self._rules().default_cpus}+'
            if memory is None:
                memory = f'{self._rules().default_mem_ratio}x'
            q = builder.add_query(
                builder.cpus_mem_query(
                    'gcp', cpus, memory, resources.region, resources.zone,
                    use_spot, resources.max_hourly_cost,
                    flags_require=_native.F_DEFAULT_FAMILY))
            plan.list_query = q
            plan.make = lambda name, res: res.copy(
                cloud=gcp, instance_type=name, accelerators=None, cpus=None,
                memory=None)
            plan.slot = builder.add_slot(query=q, **slot_common)
            return plan

        assert len(resources.accelerators) == 1, resources
        acc, acc_count = list(resources.accelerators.items())[0]
        tpu_vm = is_tpu_vm(resources)
        # 1) do accelerator rows exist at all (else: fuzzy candidates)?
        #    (common.py:657-676 on the full GCP frame)
        gate = builder.add_query(
            builder.accelerator_query(
                'gcp', acc, acc_count, None if tpu_vm else resources.cpus,
                None if tpu_vm else resources.memory, use_spot,
                resources.region, resources.zone, resources.max_hourly_cost,
                want_list=False, want_fuzzy=want_fuzzy))
        plan.gate_query = gate
        plan.fuzzy_query = gate
        acc_dict = {acc: acc_count}
        plan.make = lambda name, res: res.copy(
            cloud=gcp, instance_type=name, accelerators=acc_dict, cpus=None,
            memory=None)
        fields = dict(slot_common)
        fields.update(acc_slot_fields(acc, acc_count))
        fields['gate_query'] = gate

        if tpu_vm:
            # Fixed pseudo host 'TPU-VM' with documented vCPU / memory sizes
            # (gcp.py:775-806).
            n_cpus = 240 if 'v4' in acc else 96
            mem = 400 if 'v4' in acc else 334
            if not _fits(resources.cpus, n_cpus) or not _fits(
                    resources.memory, mem):
                return plan
            plan.explicit_instance = 'TPU-VM'
            plan.list_query = None
            fields['inst_id'] = -2
            plan.slot = builder.add_slot(**fields)
            return plan

        # 2) the host VM (gcp_catalog.py:359-393): a fixed A2/G2/A3/A4 type,
        #    or the cheapest n1 VM with enough vCPUs / memory. Region, zone,
        #    spot and max_hourly_cost deliberately do not take part.
        cpus, memory = resources.cpus, resources.memory
        if acc in rules.GCP_FIXED_HOSTS:
            group = rules.GCP_GROUP_IDS.get((acc, acc_count))
            if group is None:
                return plan
            spec = builder.cpus_mem_query('gcp', cpus, memory, group=group)
        else:
            table_cpus = rules.GCP_ACC_HOST_CPUS.get(
                acc, rules.GCP_ACC_HOST_CPUS['DEFAULT'])
            default_cpus = table_cpus.get(acc_count)
            if cpus is None and memory is None:
                if default_cpus is None:
                    # No host-VM rule for this count: the reference only gets
                    # here when accelerator rows exist (and then asserts,
                    # gcp_catalog.py:376-378); keep the gate for the fuzzy
                    # list and offer nothing.
                    return plan
                cpus = f'{default_cpus}+'
            if memory is None:
                cpu_val = int(cpus.strip('+').strip('x'))
                memory = f'{cpu_val * rules.GCP_GPU_MEMORY_CPU_RATIO}+'
            spec = builder.cpus_mem_query(
                'gcp', cpus, memory, flags_require=_native.F_HOST_FAMILY)
        host = builder.add_query(spec)
        plan.list_query = host
        fields['query'] = host
        plan.slot = builder.add_slot(**fields)
        return plan


def _fits(request: Optional[str], available: int) -> bool:
    if request is None:
        return True
    if request.endswith('+'):
        return float(request[:-1]) <= available
    return float(request) == available
