"""Microsoft Azure (placement-relevant part of sky/clouds/azure.py)."""
from typing import Any, Optional, Tuple

from skypilot_b200.catalog import rules
from skypilot_b200.clouds import cloud
from skypilot_b200.utils import registry
from skypilot_b200.utils import resources_utils


@registry.CLOUD_REGISTRY.register
class Azure(cloud.Cloud):
    """Azure: no zones in the catalog, so candidates are always per region,
    also for spot (azure.py:283-299)."""
    _REPR = 'Azure'
    _CATALOG = 'azure'

    @classmethod
    def _unsupported_features_for_resources(cls, resources: Any,
                                            region: Optional[str] = None):
        del resources, region
        return {
            cloud.CloudImplementationFeatures.CLONE_DISK_FROM_CLUSTER:
                'Migrating disk is currently not supported on Azure.',
            cloud.CloudImplementationFeatures.CUSTOM_NETWORK_TIER:
                'Custom network tier is currently not supported on Azure.',
            cloud.CloudImplementationFeatures.LOCAL_DISK:
                'Local disk is currently not supported on Azure',
        }

    def get_egress_cost(self, num_gigabytes: float) -> float:
        """Tiered $/GB out of Azure (azure.py:142-165)."""
        g = num_gigabytes
        if g > 150 * 1024:
            return 0.05 * g
        cost = 0.0
        if g >= 50 * 1024:
            cost += (g - 50 * 1024) * 0.07
            g -= 50 * 1024
        if g >= 10 * 1024:
            cost += (g - 10 * 1024) * 0.083
            g -= 10 * 1024
        if g > 1:
            cost += (g - 1) * 0.0875
        cost += 0.0
        return cost

    @classmethod
    def _disk_type(cls, disk_tier) -> str:
        tier = disk_tier
        if tier is None or tier == resources_utils.DiskTier.BEST:
            tier = resources_utils.DiskTier.HIGH if (
                tier == resources_utils.DiskTier.BEST
            ) else resources_utils.DiskTier.MEDIUM
        return {
            resources_utils.DiskTier.ULTRA: 'Disabled',
            resources_utils.DiskTier.HIGH: 'Premium_LRS',
            resources_utils.DiskTier.MEDIUM: 'Premium_LRS',
            resources_utils.DiskTier.LOW: 'Standard_LRS',
        }[tier]

    @classmethod
    def _needs_premium_disk(cls, disk_tier) -> bool:
        if disk_tier is None or disk_tier == resources_utils.DiskTier.BEST:
            return False
        if disk_tier == resources_utils.DiskTier.ULTRA:
            return False  # rejected for every instance type, see below
        return cls._disk_type(disk_tier) == 'Premium_LRS'

    @classmethod
    def check_disk_tier(cls, instance_type: Optional[str],
                        disk_tier) -> Tuple[bool, str]:
        """Premium SSDs need an S-series VM (azure.py:724-742)."""
        if disk_tier is None or disk_tier == resources_utils.DiskTier.BEST:
            return True, ''
        if disk_tier == resources_utils.DiskTier.ULTRA:
            return False, ('Azure disk_tier=ultra is not supported now. '
                           'Please use disk_tier={low, medium, high, best} '
                           'instead.')
        if (cls._disk_type(disk_tier) == 'Premium_LRS' and
                instance_type is not None and
                not rules.azure_is_s_series(instance_type)):
            return False, ('Azure premium SSDs are only supported for '
                           'S-series instances. To use disk_tier>=medium, '
                           'please make sure instance_type is specified to an '
                           'S-series instance.')
        return True, ''

    # the override below only looks at disk_tier, which the fast statement
    # path never sees (cloud.Cloud.plan_fast)
    _FAST_TEMPLATE_OK = True

    def plan_feasible(self, builder, resources: Any,
                      want_list: bool = False,
                      want_fuzzy=None) -> cloud.SlotPlan:
        if resources.disk_tier == resources_utils.DiskTier.ULTRA:
            return cloud.SlotPlan()
        return super().plan_feasible(builder, resources, want_list, want_fuzzy)
