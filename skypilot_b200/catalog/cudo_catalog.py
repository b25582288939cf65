"""Cudo catalog: the function table of sky/catalog/cudo_catalog.py, served
from the GPU-resident store (see _cloud_catalog.CloudCatalog)."""
from skypilot_b200.catalog import _cloud_catalog

_impl = _cloud_catalog.CloudCatalog('cudo', supports_zones=False)

instance_type_exists = _impl.instance_type_exists
validate_region_zone = _impl.validate_region_zone
get_hourly_cost = _impl.get_hourly_cost
get_vcpus_mem_from_instance_type = _impl.get_vcpus_mem_from_instance_type
get_default_instance_type = _impl.get_default_instance_type
get_accelerators_from_instance_type = _impl.get_accelerators_from_instance_type
get_arch_from_instance_type = _impl.get_arch_from_instance_type
get_local_disk_from_instance_type = _impl.get_local_disk_from_instance_type
get_instance_type_for_accelerator = _impl.get_instance_type_for_accelerator
get_region_zones_for_instance_type = _impl.get_region_zones_for_instance_type
list_accelerators = _impl.list_accelerators
get_image_id_from_tag = _impl.get_image_id_from_tag
is_image_tag_valid = _impl.is_image_tag_valid
