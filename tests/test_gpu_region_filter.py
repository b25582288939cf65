"""Region allow-lists of launchables (a per-region ssh_proxy_command in the
SkyPilot config; sky/resources.py:1210-1246) against records of the
unmodified reference (tests/golden/regfilter.json), and the per-region
image_id dict that takes the same device path."""
import pytest

import skypilot_b200 as sky
from tests import scenario_runner as runner
from tests import scenarios

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('scenario', scenarios.region_filter_scenarios(),
                         ids=lambda s: s['name'])
def test_region_filter_matches_reference(scenario):
    payload = runner.load_golden('regfilter')
    assert payload['catalog'] == scenarios.CATALOGS['regfilter']
    golden = next(r for r in payload['records']
                  if r['name'] == scenario['name'])
    runner.activate_catalog(payload['catalog'])
    got = runner.run_scenario(scenario)
    diffs = runner.compare(golden, got)
    assert not diffs, '\n'.join(diffs)


def test_image_id_dict_restricts_the_candidates():
    """Candidates of a request with a per-region image_id dict are the
    unrestricted candidates of those regions, in the same order."""
    runner.activate_catalog(scenarios.CATALOGS['regfilter'])
    images = {'us-west-2': 'ami-0aaaaaaaaaaaaaaaa',
              'eu-west-1': 'ami-0cccccccccccccccc'}
    free = runner.run_scenario(scenarios._single(  # pylint: disable=protected-access
        'free', cloud='aws', accelerators='V100'))
    tied = runner.run_scenario(scenarios._single(  # pylint: disable=protected-access
        'tied', cloud='aws', accelerators='V100', image_id=images))
    want = [c for c in free['candidates'][0] if c[2] in images]
    assert want and tied['candidates'][0] == want
    assert tied['plan'][0]['region'] in images
    # the host-side method agrees
    best = sky.Resources(cloud=sky.clouds.AWS(), instance_type=tied['plan'][0]
                         ['instance_type'], image_id=images)
    names = [r.name for r in best.get_valid_regions_for_launchable()]
    assert names and set(names) <= set(images)
