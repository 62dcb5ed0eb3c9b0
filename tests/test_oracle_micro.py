"""Hand-computed micro-populations: every expectation below was derived on paper from the Rust source
(file:line cited per case), not from any model in this repo.  They pin the CPU oracle (and, through
tests/test_gpu_parity.py::test_micro_cases_gpu, the CUDA path) on each branch of the update.

Notation: w = average_weight = 2/11, soc/sub/nei connectivity = 0.7/0.5/0.3, consistency 1.
"""
import numpy as np
import pytest

import motivate_b200 as mb
from common import BAD, CAR, CITY, CYCLE, DISTANT, GOOD, LOCAL, PT, WALK, tables, tiny_population
from oracle.binding import OracleSimulation

BIG = 10**6
PRM = mb.make_params()


def one_step(pop, tab, weather, changed, cls=OracleSimulation):
    sim = cls(pop, tab, PRM)
    sim.step(weather, changed)
    return sim.get_state(), sim.stats_row()


# The cases are data so the GPU test can replay them: (name, population, tables, weather, changed, expected)
def case_social_pull_and_rank_ties():
    """Two friends, Car and Walk yesterday (agent.rs:93-177, 269-316).
    Agent 0 (Car): soc = {Walk: 1*0.7/1}, sub = 0.25 each -> t = [.25,.25,.25,.95] -> norm Walk.
      avg = [(.25+1)/4, .0625, .0625, .2375]; budget ranks Car0 Walk1 PT2 Cycle3 (PT/Cycle tie -> enum order).
      cost: Local, supportiveness .5 -> all (.2+.5)/2 equal -> ranks Car0 PT1 Cycle2 Walk3.
      sums Car0 PT3 Cycle5 Walk4 -> Car.
    Agent 1 (Walk): soc = {Car: .7} -> norm Car; budget ranks Walk0 Car1 PT2 Cycle3; sums Car1 PT3 Cycle5 Walk3 -> Car.
    """
    pop = tiny_population([CAR, WALK], social=[[1], [0]], neighbours=[[], []])
    tab = tables([[.5, .5, .5, .5]], [[BIG] * 4], [[.5, .5, .5, .5]])
    return pop, tab, GOOD, 0, dict(mode=[CAR, CAR], norm=[WALK, CAR], row=[0, 1, 0, 1, 0, 0, 0, 0, 0])


def case_norm_tie_takes_last_maximum():
    """No links at all: t = desirability*0.5 = [.2,.25,.45,.45] (Subculture C, scenario.yaml:280-285).
    `max_by` keeps the LAST maximum (agent.rs:121-124) -> norm Walk, not Cycle.
    Agent was PT: avg = [.05, (.25+1)/4=.3125, .1125, .1125]; budget ranks PT0 Cycle1 Walk2 Car3
    (Cycle/Walk tie -> Cycle first).  cost Local equal -> Car0 PT1 Cycle2 Walk3.  sums Car3 PT1 Cycle3 Walk5 -> PT.
    """
    pop = tiny_population([PT], social=[[]], neighbours=[[]])
    tab = tables([[.5, .5, .5, .5]], [[BIG] * 4], [[.4, .5, .9, .9]])
    return pop, tab, GOOD, 0, dict(mode=[PT], norm=[WALK], row=[0, 1, 0, 1, 0, 0, 0, 0, 0])


def case_sum_tie_resolved_by_budget_rank():
    """Tie on the rank sum is broken by the better budget rank, not by enum order (agent.rs:304-315).
    Local commute, supportiveness [.1,.9,.1,.8] -> cost (.2+(1-s))/2 = [.55,.15,.55,.2] -> ranks PT0 Walk1 Car2 Cycle3.
    No links, desirability .5 -> t = .25 each -> norm Walk (last of equal maxima).
    Agent 0 (PT, owns car+bike): budget ranks PT0 Car1 Cycle2 Walk3 -> sums Car3 PT0 Cycle5 Walk4 -> PT.
    Agent 1 (Walk, no car, bike): budget = [0,.0625,.0625,.3125] -> ranks Walk0 PT1 Cycle2 Car3;
      allowed {PT,Cycle,Walk}: sums PT 1+0=1, Cycle 2+3=5, Walk 0+1=1 -> tie PT/Walk -> Walk has budget rank 0 -> Walk
      (taking the lower enum index would have given PT).
    """
    pop = tiny_population([PT, WALK], social=[[], []], neighbours=[[], []], car=[1, 0])
    tab = tables([[.1, .9, .1, .8]], [[BIG] * 4], [[.5, .5, .5, .5]])
    return pop, tab, GOOD, 0, dict(mode=[PT, WALK], norm=[WALK, WALK], row=[1, 2, 0, 1, 1, 0, 0, 1, 1])


def case_distant_commute_key_set():
    """DistantCommute has only Car and PT (journey_type.rs:29-32): the cost ranking runs over 2 keys and the
    choice set is {Car, PT} ∩ ownership (agent.rs:274-285).
    Agent 0 (Walk, owns car): budget ranks Walk0 Car1 PT2 Cycle3 (as above); cost Car0 PT1 -> sums Car1 PT3 -> Car.
    Agent 1 (Walk, no car): only PT remains -> PT.
    """
    pop = tiny_population([WALK, WALK], social=[[], []], neighbours=[[], []], commute=[DISTANT, DISTANT], car=[1, 0])
    tab = tables([[.5, .5, .5, .5]], [[BIG] * 4], [[.5, .5, .5, .5]])
    return pop, tab, GOOD, 0, dict(mode=[CAR, PT], norm=[WALK, WALK], row=[0, 2, 0, 2, 0, 0, 0, 0, 0])


def case_congestion_over_capacity():
    """4 residents, all drove yesterday, Car capacity 1 (neighbourhood.rs:76-85):
    modifier = 1 - (4-1)/(4-1) = 0 -> Car budget 0 (agent.rs:155-164); PT/Cycle/Walk had no users -> key absent (1.0).
    Each agent: no links, desirability [.9,.5,.5,.5]: t = [.45,.25,.25,.25] -> norm Car.
      avg = [(.45+1)/4, .0625, .0625, .0625]; budget = [0, .0625, .0625, .0625] -> ranks PT0 Cycle1 Walk2 Car3.
      cost Local equal -> Car0 PT1 Cycle2 Walk3; sums Car3 PT1 Cycle3 Walk5 -> PT.
    """
    pop = tiny_population([CAR] * 4, social=[[]] * 4, neighbours=[[]] * 4)
    tab = tables([[.5, .5, .5, .5]], [[1, BIG, BIG, BIG]], [[.9, .5, .5, .5]])
    return pop, tab, GOOD, 0, dict(mode=[PT] * 4, norm=[CAR] * 4, row=[0, 0, 0, 0, 0, 0, 0, 0, 0],
                                   cong=[[0.0, 1.0, 1.0, 1.0]])


def case_partial_congestion_value():
    """5 residents: 4 Car + 1 PT yesterday, Car capacity 2 -> 1 - (4-2)/(5-2) in f32 (neighbourhood.rs:81-83)."""
    pop = tiny_population([CAR] * 4 + [PT], social=[[]] * 5, neighbours=[[]] * 5)
    tab = tables([[.5, .5, .5, .5]], [[2, BIG, BIG, BIG]], [[.5, .5, .5, .5]])
    c = np.float32(1.0) - np.float32(2.0) / np.float32(3.0)
    return pop, tab, GOOD, 0, dict(cong=[[float(c), 1.0, 1.0, 1.0]])


def case_bad_weather_resolve():
    """Bad weather, no change (agent.rs:198-221): an agent who walked yesterday gets resolve -0.1, one who drove
    gets +0.1; sensitivity 0.5.  City commute, supportiveness .8:
      base cost = [(.3+.2)/2, .25, (.6+.2)/2=.4, (.9+.2)/2=.55].
    Agent 0 (Walk): mod = 1.5-0.1 = 1.4 -> cost [.25,.25,.56,.77] -> ranks Car0 PT1 Cycle2 Walk3.
      budget (no links, des .5): Walk0 Car1 PT2 Cycle3 -> sums Car1 PT3 Cycle5 Walk3 -> Car.
    Agent 1 (Car): budget Car0 PT1 Cycle2 Walk3, same cost order -> Car.
    With `changed` = 1 resolve is 0: orders unchanged -> same choices (checked separately below on values).
    """
    pop = tiny_population([WALK, CAR], social=[[], []], neighbours=[[], []], commute=[CITY, CITY], ws=[.5, .5])
    tab = tables([[.8, .8, .8, .8]], [[BIG] * 4], [[.5, .5, .5, .5]])
    return pop, tab, BAD, 0, dict(mode=[CAR, CAR], norm=[WALK, WALK], row=[0, 2, 0, 2, 0, 0, 0, 0, 0])


def case_bad_weather_flips_cost_order():
    """Local commute, supportiveness Car .3, PT .3, Cycle .9, Walk .9 -> base cost [.45,.45,.15,.15]: active modes cheaper.
    Good weather: cost ranks Cycle0 Walk1 Car2 PT3.  Agent was PT (habit PT), des .5, no links:
      budget PT0 Car1 Cycle2 Walk3 -> sums Car3 PT3 Cycle2 Walk4 -> Cycle.
    Bad weather, sensitivity 1.0, no change, last=PT -> resolve +.1 -> mod 2.1 -> cost [.45,.45,.315,.315]: still cheaper
      -> Cycle.  With sensitivity 3.0: mod 4.1 -> [.45,.45,.615,.615] -> ranks Car0 PT1 Cycle2 Walk3
      -> sums Car1 PT1 Cycle4 Walk6 -> tie Car/PT -> budget rank -> PT.
    """
    pop = tiny_population([PT, PT], social=[[], []], neighbours=[[], []], ws=[1.0, 3.0])
    tab = tables([[.3, .3, .9, .9]], [[BIG] * 4], [[.5, .5, .5, .5]])
    return pop, tab, BAD, 0, dict(mode=[CYCLE, PT], norm=[WALK, WALK], row=[1, 2, 0, 1, 1, 0, 0, 1, 1])


def case_duplicate_links_count_twice():
    """A friend listed several times counts several times (agent.rs:322-332; BA draws with replacement,
    social_network.rs:37-41).  Agent 0 friends [1,1,1,2]: 1 walked, 2 drove -> soc = {Walk: 3*.7/4, Car: 1*.7/4}.
    Neighbours [2]: nei = {Car: .3}.  des .5 -> sub .25.
      t[Car] = .175+.3+.25 = .725, t[Walk] = .525+.25 = .775 -> norm Walk.
    (De-duplicated lists would give Car .35+.3+.25 = .9 vs Walk .6 -> Car.)
    """
    pop = tiny_population([PT, WALK, CAR], social=[[1, 1, 1, 2], [], []], neighbours=[[2], [], []])
    tab = tables([[.5, .5, .5, .5]], [[BIG] * 4], [[.5, .5, .5, .5]])
    return pop, tab, GOOD, 0, dict(norm0=WALK)


def _resolve_population():
    pop = tiny_population([WALK], social=[[]], neighbours=[[]], ws=[.2])
    tab = tables([[.5, .5, .6, .6]], [[BIG] * 4], [[.5, .5, .5, .5]])
    return pop, tab


def case_bad_weather_unchanged_keeps_walker_walking():
    """Local commute, supportiveness [.5,.5,.6,.6] -> base cost (.2+(1-s))/2 = [.35,.35,.3,.3]; sensitivity 0.2.
    Bad again today (no change) and the agent walked: resolve -0.1 (agent.rs:205-212) -> mod = (1+.2)-.1 = 1.1 ->
    Cycle/Walk cost .33 < .35 -> cost ranks Cycle0 Walk1 Car2 PT3.
    No links, des .5: avg = [.0625,.0625,.0625,(.25+1)/4]; budget ranks Walk0 Car1 PT2 Cycle3.
    sums Car3 PT5 Cycle3 Walk1 -> Walk.  norm: t all .25 -> last maximum -> Walk.
    """
    pop, tab = _resolve_population()
    return pop, tab, BAD, 0, dict(mode=[WALK], norm=[WALK], row=[1, 1, 0, 0, 1, 0, 0, 1, 1])


def case_bad_weather_changed_has_no_resolve():
    """Same agent, but the weather CHANGED to Bad today: resolve 0.0 (agent.rs:213-215) -> mod = 1.2 -> Cycle/Walk
    cost .36 > .35 -> cost ranks Car0 PT1 Cycle2 Walk3; budget ranks as above Walk0 Car1 PT2 Cycle3.
    sums Car1 PT3 Cycle5 Walk3 -> Car.  Row: inactive mode against an active norm.
    """
    pop, tab = _resolve_population()
    return pop, tab, BAD, 1, dict(mode=[CAR], norm=[WALK], row=[0, 1, 0, 1, 0, 0, 0, 0, 0])


def case_ownership_limits_budget_and_choice():
    """Three agents, all PT yesterday, no links, des .5, Local, supportiveness [.3,.3,.9,.9]
    -> cost [.45,.45,.15,.15] -> ranks Cycle0 Walk1 Car2 PT3 (agent.rs:224-232).
    avg = [.0625, .3125, .0625, .0625]; ownership multiplies Car / Cycle by 0 or 1 (agent.rs:150-153) and removes
    them from the choice set (agent.rs:274-285).
    Agent 0 (car, bike):    budget ranks PT0 Car1 Cycle2 Walk3 -> sums Car3 PT3 Cycle2 Walk4 -> Cycle.
    Agent 1 (car, no bike): b = [.0625,.3125,0,.0625] -> PT0 Car1 Walk2 Cycle3; allowed {Car,PT,Walk}:
                            sums Car3 PT3 Walk3 -> three-way tie -> best budget rank -> PT.
    Agent 2 (neither):      b = [0,.3125,0,.0625] -> PT0 Walk1 Car2 Cycle3; allowed {PT,Walk}: PT3 Walk2 -> Walk.
    """
    pop = tiny_population([PT, PT, PT], social=[[]] * 3, neighbours=[[]] * 3, car=[1, 1, 0], bike=[1, 0, 0])
    tab = tables([[.3, .3, .9, .9]], [[BIG] * 4], [[.5, .5, .5, .5]])
    return pop, tab, GOOD, 0, dict(mode=[CYCLE, PT, WALK], norm=[WALK] * 3, row=[2, 3, 0, 1, 2, 0, 0, 2, 2])


def case_two_neighbourhoods_two_subcultures():
    """Congestion and the statistics columns are per neighbourhood / per subculture (neighbourhood.rs:57-89,
    statistics.rs:62-93).  Everybody drove yesterday.  Neighbourhood 0: Car capacity 1 for 2 residents -> modifier
    1-(2-1)/(2-1) = 0, supportiveness [.5,.5,.5,.9] -> cost [.35,.35,.35,.15] -> ranks Walk0 Car1 PT2 Cycle3.
    Neighbourhood 1: uncongested, supportiveness .5 -> cost ranks Car0 PT1 Cycle2 Walk3.
    Subculture 0 des .5 -> sub .25; subculture 1 des [.1,.1,.1,.9] -> sub [.05,.05,.05,.45].
    A0 (n0,s0): b = [0,.0625,.0625,.0625] -> PT0 Cycle1 Walk2 Car3 -> sums Car4 PT2 Cycle4 Walk2 -> tie -> PT.
    A1 (n0,s1): avg = [.2625,.0125,.0125,.1125], b = [0,.0125,.0125,.1125] -> Walk0 PT1 Cycle2 Car3 -> Walk sum 0 -> Walk.
    A2 (n1,s0), A3 (n1,s1): Car has budget rank 0 and cost rank 0 -> Car.
    """
    pop = tiny_population([CAR] * 4, social=[[]] * 4, neighbours=[[]] * 4, nbhd=[0, 0, 1, 1], sub=[0, 1, 0, 1])
    tab = tables([[.5, .5, .5, .9], [.5, .5, .5, .5]], [[1, BIG, BIG, BIG], [BIG] * 4], [[.5, .5, .5, .5], [.1, .1, .1, .9]])
    return pop, tab, GOOD, 0, dict(mode=[PT, WALK, CAR, CAR], norm=[WALK] * 4, row=[1, 4, 0, 3, 1, 0, 0, 0, 1, 1, 0],
                                   cong=[[0.0, 1.0, 1.0, 1.0], [1.0, 1.0, 1.0, 1.0]])


CASES = [case_social_pull_and_rank_ties, case_norm_tie_takes_last_maximum, case_sum_tie_resolved_by_budget_rank,
         case_distant_commute_key_set, case_congestion_over_capacity, case_partial_congestion_value,
         case_bad_weather_resolve, case_bad_weather_flips_cost_order, case_duplicate_links_count_twice,
         case_bad_weather_unchanged_keeps_walker_walking, case_bad_weather_changed_has_no_resolve,
         case_ownership_limits_budget_and_choice, case_two_neighbourhoods_two_subcultures]


def check_case(case, cls):
    pop, tab, weather, changed, exp = case()
    st, row = one_step(pop, tab, weather, changed, cls)
    if "mode" in exp:
        assert st["current_mode"].tolist() == exp["mode"], case.__name__
    if "norm" in exp:
        assert st["norm"].tolist() == exp["norm"], case.__name__
    if "norm0" in exp:
        assert st["norm"][0] == exp["norm0"]
    if "row" in exp:
        assert row.tolist() == exp["row"], case.__name__
    if "cong" in exp:
        assert np.array_equal(st["congestion"], np.array(exp["cong"], np.float32)), case.__name__
    # last_mode is yesterday's current_mode (agent.rs:245)
    assert np.array_equal(st["last_mode"], pop["current_mode"])


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.__name__)
def test_micro_case_oracle(case):
    check_case(case, OracleSimulation)


def test_habit_update_values():
    """agent.rs:244-264: h[m] *= (1-w); h[last] += w, in f32, multiply before add."""
    w = np.float32(2.0) / (np.float32(10) + np.float32(1.0))
    keep = np.float32(1.0) - w
    pop = tiny_population([CYCLE], social=[[]], neighbours=[[]], habit=[[0.25, 0.0, 0.5, 0.125]])
    tab = tables([[.5] * 4], [[BIG] * 4], [[.5] * 4])
    st, _ = one_step(pop, tab, GOOD, 0)
    exp = np.array([np.float32(0.25) * keep, np.float32(0.0), np.float32(0.5) * keep + w, np.float32(0.125) * keep], np.float32)
    assert np.array_equal(st["habit"][0], exp)


def test_weekday_gate_prev_weather_and_weekend_intervention():
    """simulation.rs:101-136: rows only for day%7<5; `change` compares with the previous WEEKDAY's weather (:131);
    an intervention dated on a weekend still applies (:103-105 precede the gate)."""
    pop = tiny_population([PT, PT], social=[[], []], neighbours=[[], []], ws=[3.0, 3.0])
    tab = tables([[.3, .3, .9, .9]], [[BIG] * 4], [[.5, .5, .5, .5]])
    # days:      0     1     2    3    4    5(sat) 6(sun) 7     8
    weather = [GOOD, GOOD, BAD, BAD, BAD, GOOD,  GOOD,  BAD,  GOOD]
    sim = OracleSimulation(pop, tab, PRM)
    days, rows = sim.run(np.array(weather, np.uint8))
    assert days.tolist() == [0, 1, 2, 3, 4, 7, 8]
    assert mb.run_rows(9) == 7 and mb.run_rows(730) == 522 and mb.run_rows(1825) == 1305
    # day 7 is Bad and the previous weekday (4) was Bad -> no change although days 5,6 were Good.
    # Replay by hand with explicit steps and compare.
    sim2 = OracleSimulation(pop, tab, PRM)
    prev = weather[0]
    for d in (1, 2, 3, 4, 7, 8):
        sim2.step(weather[d], prev != weather[d])
        prev = weather[d]
    a, b = sim.get_state(), sim2.get_state()
    assert np.array_equal(a["current_mode"], b["current_mode"]) and np.array_equal(a["habit"], b["habit"])
    # weekend intervention (day 6): car taken away -> from day 7 the agent cannot choose Car
    tab2 = tab.copy()
    iv = mb.InterventionData(day=6, tables=tab2, owns_car=np.array([0, 0], np.uint8), owns_bike=np.array([0, 0], np.uint8))
    sim3 = OracleSimulation(pop, tab, PRM)
    days3, rows3 = sim3.run(np.array(weather, np.uint8), intervention=iv)
    assert set(sim3.get_state()["current_mode"].tolist()) <= {PT, WALK}
    assert np.array_equal(rows3[:5], rows[:5])       # identical until the intervention
