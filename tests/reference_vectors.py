"""Known-answer vectors of the reference's own unit tests, as data.

Source: tests/unit_tests/test_catalog.py of skypilot-org/skypilot @ 7808630
  * AZ frame + 7 cases                      (:89-142)
  * no-AZ frame + 6 cases                   (:145-183)
  * hourly cost returns a Python float      (:186-220)
  * local-disk frame, 9 selection cases     (:260-344), 3 set cases (:347-366)
and tests/unit_tests/test_resources.py `should_be_blocked_by` (:1207-1293),
tests/unit_tests/test_dag.py `is_chain` (:93-149).
Used twice: against the pandas oracle (CPU) and against the CUDA path (GPU).
"""
import pandas as pd

NAN = float('nan')


def az_frame() -> pd.DataFrame:
    rows = [('a', 4, 16, 1.0, 'us-west1', 'us-west1-a'),
            ('b', 8, 32, 2.0, 'us-west1', 'us-west1-b'),
            ('c', 8, 32, 5.0, 'us-west1', 'us-west1-a'),
            ('d', 8, 32, 3.0, 'asia-southeast1', 'asia-southeast1-a')]
    return pd.DataFrame(rows, columns=['InstanceType', 'vCPUs', 'MemoryGiB',
                                       'Price', 'Region', 'AvailabilityZone'])


# (cpus, memory, region, zone, expected instance type)
AZ_CASES = [
    ('4', '16', None, None, 'a'),
    ('4+', '16+', None, None, 'a'),
    ('16', '128', None, None, None),
    ('1+', None, 'asia-southeast1', None, 'd'),
    ('1+', None, 'us-west1', 'us-west1-b', 'b'),
    ('1+', None, 'us-west1', 'us-west1-c', None),
    # b is cheaper but only offered in us-west1-b => c
    ('8', '32', 'us-west1', 'us-west1-a', 'c'),
]


def no_az_frame() -> pd.DataFrame:
    rows = [('a', 4, 16, 1.0, 'us-east1'), ('b', 4, 16, 3.0,
                                             'asia-southeast1'),
            ('c', 8, 32, 5.0, 'asia-southeast1')]
    return pd.DataFrame(
        rows, columns=['InstanceType', 'vCPUs', 'MemoryGiB', 'Price',
                       'Region'])


# (cpus, memory, region, expected)
NO_AZ_CASES = [
    ('4', '16', None, 'a'),
    ('1+', None, None, 'a'),
    ('8+', '32+', None, 'c'),
    ('16', '128', None, None),
    ('1+', None, 'asia-southeast1', 'b'),
    ('1+', None, 'europe-west1', None),
]


def price_frame() -> pd.DataFrame:
    return pd.DataFrame([{
        'InstanceType': 'test-instance', 'Price': 1.5, 'SpotPrice': 0.5,
        'Region': 'us-west1', 'AvailabilityZone': 'us-west1-a'
    }])


def local_disk_frame() -> pd.DataFrame:
    rows = [
        ('m5.large', 2, 8, 0.10, NAN, NAN, NAN, False),
        ('i3.large', 2, 15.25, 0.16, 'ssd', 475.0, 1, True),
        ('i3.2xlarge', 8, 61, 0.62, 'ssd', 950.0, 2, True),
        ('d2.large', 2, 15.25, 0.14, 'ssd', 500.0, 1, False),
    ]
    df = pd.DataFrame(rows, columns=[
        'InstanceType', 'vCPUs', 'MemoryGiB', 'Price', 'LocalDiskType',
        'LocalDiskSize', 'LocalDiskCount', 'NVMeSupported'
    ])
    df['Region'] = 'us-east-1'
    df['AvailabilityZone'] = 'us-east-1a'
    return df


# (local_disk, expected cheapest instance with cpus='1+')
LOCAL_DISK_CASES = [
    ('nvme:500+', 'i3.2xlarge'),
    ('nvme:1500+', 'i3.2xlarge'),
    ('nvme:100+', 'i3.large'),
    ('ssd:400+', 'd2.large'),
    ('nvme:475', 'i3.large'),
    ('nvme:300', None),
    ('ssd:500', 'd2.large'),
    (None, 'm5.large'),
    ('nvme:5000+', None),
]

# (local_disk, instance types that survive the filter)
LOCAL_DISK_SETS = [
    ('nvme:100+', ['i3.2xlarge', 'i3.large']),
    ('ssd:100+', ['d2.large', 'i3.2xlarge', 'i3.large']),
    (None, ['d2.large', 'i3.2xlarge', 'i3.large', 'm5.large']),
]

_P3 = dict(cloud='aws', instance_type='p3.2xlarge', region='us-west-2',
           zone='us-west-2a', accelerators={'V100': 1})
# (candidate kwargs, blocked kwargs, expected)
BLOCKED_CASES = [
    (dict(_P3, use_spot=True), dict(_P3, use_spot=True), True),
    (dict(_P3, use_spot=True), dict(_P3, use_spot=False), False),
    (dict(_P3), dict(cloud='gcp', instance_type='g2-standard-4',
                     region='us-east4', zone='us-east4-c',
                     accelerators={'L4': 1}), False),
    (dict(_P3, use_spot=True),
     dict(_P3, instance_type='p3.8xlarge', use_spot=True), False),
]

# (edges over tasks 0..n-1, n, expected is_chain)
CHAIN_CASES = [
    ([], 1, True),
    ([(0, 1), (1, 2)], 3, True),
    ([(0, 1), (0, 2)], 3, False),          # fork
    ([(0, 2), (1, 2)], 3, False),          # join
    ([(0, 1), (0, 2), (1, 3), (2, 3)], 4, False),  # diamond
    ([], 2, False),                        # two isolated tasks
    ([(0, 1)], 3, False),                  # a chain plus an isolated task
]
