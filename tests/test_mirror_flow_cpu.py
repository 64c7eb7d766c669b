"""Host logic of the drop-in `Country` end to end on CPU, with the engine replaced by the oracle-backed test double
(tests/fake_sim.py): person/canonical permutation, seeding, read-back properties, daily log lines, scenario clock
hooks -> refresh, pickling and resuming.  The same flows run against the real CUDA engine in test_gpu_dropin.py."""
from datetime import datetime

import numpy as np
import pytest

import dropin_flow
from fake_sim import FakeSimulation


@pytest.fixture()
def fake_engine(monkeypatch):
    from abm_b200 import engine
    monkeypatch.setattr(engine, 'Simulation', FakeSimulation)
    FakeSimulation.instances.clear()
    yield FakeSimulation


def test_unmitigated_flow_with_pickle_resume(tmp_path, fake_engine):
    m, ora, rows = dropin_flow.run_flow(tmp_path, 'UnmitigatedScenario', days=8, n_agents=6000, pickle_at=6 * 4 + 3)
    assert rows[-1]['infected_count'] > 150
    assert m.epidemic_state.dtype == np.float64
    assert m.scheduler.real_time == datetime(2020, 9, 9, 0) and m.scheduler.steps == 48


def test_lockdown_flow(tmp_path, fake_engine):
    m, ora, rows = dropin_flow.run_flow(tmp_path, 'LockdownGreatestMobilityScenario', days=5, n_agents=6000)
    assert m.params.lockdown and not m.lockdown                     # quirk Q1


def test_school_scenario_clock_hooks_refresh_the_device(tmp_path, fake_engine):
    m, ora, rows = dropin_flow.run_flow(tmp_path, 'AcceleratedGovernmentOpenSchoolsScenario', days=4,
                                        start=datetime(2020, 9, 29, 0), n_agents=6000, school_columns=True)
    assert fake_engine.instances[-1].n_refresh >= 1
    assert m.active_school_phases[-1] == m.max_school_phase        # one phase per sub-step of the 1st (quirk Q10)
    assert (m.economic_activity_location_ids >= max(m.params.HOUSEHOLD_ID_TO_NAME) + 1).sum() > 100


def test_real_time_cannot_be_set_while_running(tmp_path, fake_engine):
    m, _, _ = dropin_flow.run_flow(tmp_path, 'UnmitigatedScenario', days=1, n_agents=3000)
    with pytest.raises(RuntimeError):
        m.scheduler.real_time = datetime(2021, 1, 1)
    groups = m.get_location_person_ids()
    assert sum(len(v) for v in groups.values()) == m.person_ids.shape[0]
    assert m.scheduler.get_active_ids().shape[0] > 0


@pytest.mark.parametrize('scenario', ['DynamicPhase1GovernmentOpenSchoolsScenario', 'Phase1GovernmentOpenSchoolsScenario'])
def test_other_school_scenarios_run(tmp_path, fake_engine, scenario):
    """The monthly clock hooks of the two other government school scenarios (base_model.py:809-834): the dynamic one
    asks for the safe / unsafe districts on the 1st of the month (quirk Q10: it gets rates back, so nobody matches)."""
    start = datetime(2020, 12, 30, 0) if scenario.startswith('Phase1') else datetime(2020, 9, 29, 0)
    m, ora, rows = dropin_flow.run_flow(tmp_path, scenario, days=4, start=start, n_agents=6000, school_columns=True)
    safe, unsafe = m.get_safe_and_unsafe_districts()
    assert safe.shape[0] + unsafe.shape[0] == 60 and safe.max() <= unsafe.min() if unsafe.size else True
    if scenario.startswith('Phase1'):
        assert m.active_school_phases[:3] == [1, 2, 4]            # appended on every sub-step of 2021-01-01 (Q10)


def test_open_mining_flow_with_camps(tmp_path, fake_engine):
    """OpenMiningScenario (scenario_models.py:224-246): miners live and work in the camp of their district -- one
    household of tens of people (beyond the 32-member household tile), exempt from the lockdown, never travelling."""
    m, ora, rows = dropin_flow.run_flow(tmp_path, 'OpenMiningScenario', days=8, n_agents=12000, mining_columns=True,
                                        mining_share=0.6)
    names = m.params.HOUSEHOLD_ID_TO_NAME
    camp_ids = [i for i, n in names.items() if n.startswith('mining_')]
    assert len(camp_ids) == 12 and min(camp_ids) == max(names) - 11           # sorted after every 'h_...' household
    in_camp = np.isin(m.household_ids, camp_ids)
    size = np.bincount(m.household_ids[in_camp] - min(camp_ids))
    assert size.max() > 32 and rows[-1]['infected_count'] > 100
    assert not m.district_mover[in_camp].any()
    assert np.array_equal(m.economic_activity_location_ids[in_camp], m.household_ids[in_camp])
    # miners keep their full weekday movement probability; their locked-down neighbours do not
    assert (m.economic_status_weekday_movement_probability[in_camp] == 1.0).all()
    assert m.economic_status_weekday_movement_probability[~in_camp].mean() < 0.9
