"""Lambda Cloud (placement-relevant part of sky/clouds/lambda_cloud.py)."""
from typing import Any, Optional

from skypilot_b200.clouds import cloud
from skypilot_b200.utils import registry


@registry.CLOUD_REGISTRY.register
class Lambda(cloud.Cloud):
    """Lambda: no zones, no spot (`regions_with_offering` returns [] for
    spot requests, lambda_cloud.py:84-86)."""
    _REPR = 'Lambda'
    _CATALOG = 'lambda'

    @classmethod
    def _unsupported_features_for_resources(cls, resources: Any,
                                            region: Optional[str] = None):
        del resources, region
        features = cloud.CloudImplementationFeatures
        return {
            features.STOP: 'Lambda cloud does not support stopping VMs.',
            features.CLONE_DISK_FROM_CLUSTER:
                'Migrating disk is currently not supported on Lambda.',
            features.SPOT_INSTANCE:
                'Spot instances are not supported in Lambda.',
            features.IMAGE_ID:
                'Specifying image ID is not supported in Lambda Cloud.',
            features.CUSTOM_DISK_TIER:
                'Custom disk tiers are not supported in Lambda Cloud.',
            features.CUSTOM_NETWORK_TIER:
                'Custom network tier is not supported in Lambda Cloud.',
            features.HOST_CONTROLLERS:
                'Host controllers are not supported in Lambda Cloud.',
            features.HIGH_AVAILABILITY_CONTROLLERS:
                'High availability controllers are not supported on Lambda.',
            features.CUSTOM_MULTI_NETWORK:
                'Customized multiple network interfaces are not supported on '
                'Lambda.',
            features.LOCAL_DISK: 'Local disk is not supported on Lambda',
        }

    @classmethod
    def regions_with_offering(cls, instance_type, accelerators, use_spot,
                              region, zone, resources=None):
        assert zone is None, 'Lambda does not support zones.'
        if use_spot:
            return []
        return super().regions_with_offering(instance_type, accelerators,
                                             use_spot, region, zone, resources)
