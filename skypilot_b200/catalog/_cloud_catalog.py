"""The per-cloud catalog function table, backed by the GPU-resident store.

One `CloudCatalog` instance per cloud provides the names that the
reference's `sky/catalog/__init__.py:19-53` dispatcher looks up on
`sky.catalog.<cloud>_catalog` (instance_type_exists, validate_region_zone,
get_hourly_cost, get_default_instance_type, get_instance_type_for_accelerator,
get_region_zones_for_instance_type, ...). The thin `<cloud>_catalog.py`
modules re-export the bound methods so a drop-in import keeps working.
"""
from typing import Dict, List, Optional, Tuple, Union

from skypilot_b200 import _native
from skypilot_b200.catalog import common
from skypilot_b200.utils import resources_utils
from skypilot_b200.catalog import rules as rules_lib


class CloudCatalog:

    def __init__(self, cloud: str, supports_zones: bool = True):
        self.cloud = cloud
        self.rules = rules_lib.rules_for(cloud)
        self.supports_zones = supports_zones

    # -- plumbing -------------------------------------------------------------
    def _view(self) -> common.CatalogView:
        from skypilot_b200 import catalog  # pylint: disable=import-outside-toplevel
        return catalog.view(self.cloud)

    def _no_zone(self, zone: Optional[str]) -> None:
        if zone is not None and not self.supports_zones:
            raise ValueError(f'{self.cloud.capitalize()} does not support '
                             'zones.')

    # -- the function table ---------------------------------------------------
    def instance_type_exists(self, instance_type: str) -> bool:
        return common.instance_type_exists_impl(self._view(), instance_type)

    def validate_region_zone(
            self, region: Optional[str],
            zone: Optional[str]) -> Tuple[Optional[str], Optional[str]]:
        self._no_zone(zone)
        return common.validate_region_zone_impl(self.cloud, self._view(),
                                                region, zone)

    def get_hourly_cost(self, instance_type: str, use_spot: bool = False,
                        region: Optional[str] = None,
                        zone: Optional[str] = None) -> float:
        self._no_zone(zone)
        return common.get_hourly_cost_impl(self._view(), instance_type,
                                           use_spot, region, zone)

    def get_vcpus_mem_from_instance_type(
            self,
            instance_type: str) -> Tuple[Optional[float], Optional[float]]:
        return common.get_vcpus_mem_from_instance_type_impl(
            self._view(), instance_type)

    def get_accelerators_from_instance_type(
            self,
            instance_type: str) -> Optional[Dict[str, Union[int, float]]]:
        return common.get_accelerators_from_instance_type_impl(
            self._view(), instance_type)

    def get_arch_from_instance_type(self, instance_type: str) -> Optional[str]:
        return common.get_arch_from_instance_type_impl(self._view(),
                                                       instance_type)

    def get_local_disk_from_instance_type(
            self, instance_type: str) -> Optional[str]:
        return common.get_local_disk_from_instance_type_impl(
            self._view(), instance_type)

    def _default_view(self, disk_tier, local_disk) -> common.CatalogView:
        flags = _native.F_DEFAULT_FAMILY
        view = self._view().restrict(flags_require=flags)
        if self.rules.supports_local_disk:
            view = common.filter_with_local_disk(view, local_disk)
        if self.rules.premium_disk is not None and disk_tier is not None:
            if not isinstance(disk_tier, resources_utils.DiskTier):
                # a caller outside this package (the reference's own clouds
                # when only the catalog function table is swapped,
                # INTEGRATION.md level 2) passes its own enum
                disk_tier = resources_utils.DiskTier(
                    getattr(disk_tier, 'value', disk_tier))
            # e.g. Azure: premium SSD tiers need an S-series VM
            # (sky/catalog/azure_catalog.py:108-112)
            from skypilot_b200.utils import registry  # pylint: disable=import-outside-toplevel
            cloud_cls = registry.CLOUD_REGISTRY.get(self.cloud)
            if cloud_cls is not None and cloud_cls._needs_premium_disk(  # pylint: disable=protected-access
                    disk_tier):
                view = view.restrict(flags_require=_native.F_PREMIUM_DISK)
        return view

    def get_default_instance_type(self, cpus: Optional[str] = None,
                                  memory: Optional[str] = None,
                                  disk_tier=None,
                                  local_disk: Optional[str] = None,
                                  region: Optional[str] = None,
                                  zone: Optional[str] = None,
                                  use_spot: bool = False,
                                  max_hourly_cost: Optional[float] = None
                                  ) -> Optional[str]:
        """Cheapest instance of the default families (e.g.
        sky/catalog/aws_catalog.py:249-274)."""
        if (cpus is None and
                (memory is None or self.rules.default_cpus_always) and
                self.rules.default_cpus is not None):
            cpus = (f'{self.rules.default_cpus}' if self.rules.default_cpus_exact
                    else f'{self.rules.default_cpus}+')
        if memory is None and self.rules.default_memory is not None:
            memory = self.rules.default_memory
        elif memory is None and self.rules.default_mem_ratio is not None:
            memory = f'{self.rules.default_mem_ratio}x'
        view = self._default_view(disk_tier, local_disk)
        return common.get_instance_type_for_cpus_mem_impl(
            view, cpus, memory, region, zone, use_spot, max_hourly_cost)

    def get_instance_type_for_accelerator(
            self, acc_name: str, acc_count: Union[int, float],
            cpus: Optional[str] = None, memory: Optional[str] = None,
            use_spot: bool = False, local_disk: Optional[str] = None,
            region: Optional[str] = None, zone: Optional[str] = None,
            max_hourly_cost: Optional[float] = None
    ) -> Tuple[Optional[List[str]], List[str]]:
        """(instance types sorted by price, fuzzy candidates), e.g.
        sky/catalog/aws_catalog.py:292-319."""
        self._no_zone(zone)
        view = self._view()
        if self.rules.supports_local_disk:
            view = common.filter_with_local_disk(view, local_disk)
        return common.get_instance_type_for_accelerator_impl(
            view, acc_name, acc_count, cpus, memory, use_spot, region, zone,
            max_hourly_cost)

    def get_region_zones_for_instance_type(self, instance_type: str,
                                           use_spot: bool):
        return common.get_region_zones_for_instance_type_impl(
            self._view(), instance_type, use_spot,
            us_first=self.rules.us_regions_first)

    def list_accelerators(self, gpus_only: bool,
                          name_filter: Optional[str] = None,
                          region_filter: Optional[str] = None,
                          quantity_filter: Optional[int] = None,
                          case_sensitive: bool = True,
                          all_regions: bool = False,
                          require_price: bool = True):
        """Instance types offering accelerators, grouped by accelerator name
        (`list_accelerators` of sky/catalog/aws_catalog.py:339-352 and the
        other single-table clouds; common.py:697-790)."""
        del require_price  # unused, as in the reference
        from skypilot_b200.catalog import listing  # pylint: disable=import-outside-toplevel
        return listing.generic_listing(self.cloud, self._view(), gpus_only,
                                       name_filter, region_filter,
                                       quantity_filter, case_sensitive,
                                       all_regions)

    # -- images.csv ------------------------------------------------------
    def _images(self):
        import pandas as pd  # pylint: disable=import-outside-toplevel
        from skypilot_b200 import catalog  # pylint: disable=import-outside-toplevel
        df = catalog.get_store().images.get(self.cloud)
        if df is None:
            df = pd.DataFrame(columns=['Tag', 'Region', 'ImageId'])
        return df

    def get_image_id_from_tag(self, tag: str,
                              region: Optional[str] = None) -> Optional[str]:
        """sky/catalog/<cloud>_catalog.get_image_id_from_tag (the reference
        re-downloads images.csv once when the tag is unknown; here the table
        is whatever the catalog directory holds)."""
        return common.get_image_id_from_tag_impl(self._images(), tag, region)

    def is_image_tag_valid(self, tag: str, region: Optional[str]) -> bool:
        return common.is_image_tag_valid_impl(self._images(), tag, region)
