"""Request alternatives from a task YAML's `resources:` block --
Resources.from_yaml_config (`any_of`, `ordered`, several accelerators,
accelerators by memory size / manufacturer) -- against records of the
unmodified reference (tests/golden/altres.json)."""
import pytest

from tests import scenario_runner as runner
from tests import scenarios

pytestmark = pytest.mark.gpu

_payload = None


def _golden():
    global _payload
    if _payload is None:
        _payload = runner.load_golden('altres')
        assert _payload['catalog'] == scenarios.CATALOGS['altres']
    return _payload


@pytest.mark.parametrize('scenario', scenarios.alternatives_scenarios(),
                         ids=lambda s: s['name'])
def test_alternatives_match_reference(scenario):
    payload = _golden()
    golden = next(r for r in payload['records']
                  if r['name'] == scenario['name'])
    runner.activate_catalog(payload['catalog'])
    got = runner.run_scenario(scenario)
    # a set of Resources iterates in address order (identity hashes), in the
    # reference and here: candidate tables of such tasks compare as sets
    diffs = runner.compare(golden, got, unordered_candidates=True)
    assert not diffs, '\n'.join(diffs)
