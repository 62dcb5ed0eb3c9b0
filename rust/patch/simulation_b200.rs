//! simulation_b200.rs — drop-in for the reference crate (`src/simulation_b200.rs`): the weekday loop of
//! `simulation::run` (`src/simulation.rs:101-136`) handed to libmotivate_b200 through `motivate-b200-sys`.
//!
//! How to apply (see README.md in this directory for the exact diff):
//!   Cargo.toml   [dependencies] motivate-b200-sys = { path = ".../rust/motivate-b200-sys" }
//!   main.rs      extern crate motivate_b200_sys;   mod simulation_b200;
//!   simulation.rs: after `link_agents(..)` (:78) and the CSV header (:92) replace lines 94-136 by
//!                  `simulation_b200::run_on_gpu(&id, &scenario, &residents, weather_pattern, total_years, gpu, &mut file)?;`
//!
//! Everything the reference keeps stays as it is: `Parameters`, `Scenario`, `Intervention`, `--generate`, the
//! network / agent YAML files, `intervene`'s random sampling (it runs here on COPIES, up front, see
//! `resolve_intervention`), the CSV format.  Not compiled in the motivate-b200 build image (no cargo/rustc).

use std::cell::RefCell;
use std::collections::HashMap;
use std::io;
use std::io::Write;
use std::rc::Rc;

use rand::seq::sample_slice_ref;
use rand::thread_rng;

use agent::Agent;
use journey_type::JourneyType;
use motivate_b200_sys as mvb;
use scenario::Scenario;
use transport_mode::TransportMode;
use weather::Weather;

/// `TransportMode as u8` in the enum's declaration order (`src/transport_mode.rs:3-8`).
const MODES: [TransportMode; 4] = [TransportMode::Car, TransportMode::PublicTransport, TransportMode::Cycle, TransportMode::Walk];

fn mode_code(m: TransportMode) -> u8 {
    match m {
        TransportMode::Car => 0,
        TransportMode::PublicTransport => 1,
        TransportMode::Cycle => 2,
        TransportMode::Walk => 3,
    }
}

fn commute_code(j: JourneyType) -> u8 {
    match j {
        JourneyType::LocalCommute => 0,
        JourneyType::CityCommute => 1,
        JourneyType::DistantCommute => 2,
    }
}

/// The linked agent graph (`Vec<Rc<RefCell<Agent>>>`, `src/simulation.rs:64-78`) as structure-of-arrays.
/// Agent id = position in `residents`; subculture / neighbourhood index = position in
/// `scenario.subcultures` / `scenario.neighbourhoods` (the CSV column order, `:207-211`); the two link lists
/// become CSR graphs WITH duplicates kept, in list order (`count_in_subgroup` counts a friend listed twice
/// twice, `src/agent.rs:322-332`).
pub fn flatten(residents: &[Rc<RefCell<Agent>>], scenario: &Scenario) -> mvb::Population {
    let n = residents.len();
    let index_of: HashMap<*const RefCell<Agent>, u32> = residents
        .iter()
        .enumerate()
        .map(|(i, a)| (&**a as *const RefCell<Agent>, i as u32))
        .collect();
    let sub_index: HashMap<&str, u8> = scenario.subcultures.iter().enumerate().map(|(i, s)| (s.id.as_str(), i as u8)).collect();
    let nbhd_index: HashMap<&str, u16> = scenario.neighbourhoods.iter().enumerate().map(|(i, s)| (s.id.as_str(), i as u16)).collect();

    let mut p = mvb::Population::default();
    p.habit = vec![0f32; 4 * n];
    p.social_rowptr.push(0);
    p.neighbour_rowptr.push(0);
    for (i, cell) in residents.iter().enumerate() {
        let a = cell.borrow();
        p.subculture.push(*sub_index.get(a.subculture.id.as_str()).expect("Subculture not found"));
        p.neighbourhood.push(*nbhd_index.get(a.neighbourhood.id.as_str()).expect("Neighbourhood not found"));
        p.commute.push(commute_code(a.commute_length));
        p.owns_car.push(a.owns_car as u8);
        p.owns_bike.push(a.owns_bike as u8);
        p.weather_sensitivity.push(a.weather_sensitivity);
        // per-agent weights: every population the reference generates has uniform ones
        // (agent_generation.rs:226,243-246) but they are fields of config/agents/N.yaml, so pass them through
        p.consistency.push(a.consistency);
        p.social_connectivity.push(a.social_connectivity);
        p.subculture_connectivity.push(a.subculture_connectivity);
        p.neighbourhood_connectivity.push(a.neighbourhood_connectivity);
        p.average_weight.push(a.average_weight);
        for (m, v) in a.habit.iter() {
            p.habit[4 * i + mode_code(*m) as usize] = *v; // absent key == 0.0 (agent.rs:244-264 restated densely)
        }
        p.current_mode.push(mode_code(a.current_mode));
        p.norm.push(mode_code(a.norm));
        for f in a.social_network.iter() {
            p.social_col.push(index_of[&(&**f as *const RefCell<Agent>)]);
        }
        p.social_rowptr.push(p.social_col.len() as u64);
        for f in a.neighbours.iter() {
            p.neighbour_col.push(index_of[&(&**f as *const RefCell<Agent>)]);
        }
        p.neighbour_rowptr.push(p.neighbour_col.len() as u64);
    }
    p
}

/// supportiveness[H][4], capacity[H][4], desirability[S][4] in `TransportMode` order.  All four modes must be
/// present in every map (true of the shipped scenario; `capacity` already panics otherwise,
/// `src/neighbourhood.rs:77`): a missing key would change the key set of the budget (SURVEY.md §8 a-K).
pub fn tables_of(scenario: &Scenario) -> mvb::Tables {
    let mut t = mvb::Tables::default();
    t.n_subcultures = scenario.subcultures.len() as u32;
    t.n_neighbourhoods = scenario.neighbourhoods.len() as u32;
    for nb in scenario.neighbourhoods.iter() {
        for m in MODES.iter() {
            t.supportiveness.push(*nb.supportiveness.borrow().get(m).expect("supportiveness needs all four transport modes"));
            t.capacity.push(*nb.capacity.borrow().get(m).expect("capacity needs all four transport modes"));
        }
    }
    for sc in scenario.subcultures.iter() {
        for m in MODES.iter() {
            t.desirability.push(*sc.desirability.borrow().get(m).expect("desirability needs all four transport modes"));
        }
    }
    t
}

/// `intervene` (`src/simulation.rs:304-464`) run once, up front, on COPIES: returns the tables and ownership
/// bits that hold from `scenario.intervention.day` on.  Same arithmetic (f32 `+` for supportiveness and
/// desirability, i64 add clamped to [0, u32::MAX] for capacity), same sampling (`thread_rng` +
/// `sample_slice_ref` over the current non-owners / owners), same panics.  Returns None when the
/// intervention can never fire (`day == 0`: the loop starts at day 1, `:101-103`).
pub fn resolve_intervention(scenario: &Scenario, flat: &mvb::Population) -> Option<mvb::Intervention> {
    let iv = &scenario.intervention;
    if iv.day == 0 {
        return None;
    }
    let mut t = tables_of(scenario);
    for change in iv.neighbourhood_changes.iter() {
        let h = scenario.neighbourhoods.iter().position(|nb| nb.id == change.id)
            .expect("A neighbourhood in your intervention was not found");
        for (k, m) in MODES.iter().enumerate() {
            if let Some(d) = change.increase_in_supportiveness.get(m) {
                t.supportiveness[4 * h + k] = t.supportiveness[4 * h + k] + *d; // :320-325
            }
            if let Some(d) = change.increase_in_capacity.get(m) {
                let v = t.capacity[4 * h + k] as i64 + *d; // :328-351
                t.capacity[4 * h + k] = if v < 0 { 0 } else if v > u32::max_value() as i64 { u32::max_value() } else { v as u32 };
            }
        }
    }
    for change in iv.subculture_changes.iter() {
        let s = scenario.subcultures.iter().position(|sc| sc.id == change.id)
            .expect("A subculture in your intervention was not found");
        for (k, m) in MODES.iter().enumerate() {
            if let Some(d) = change.increase_in_desirability.get(m) {
                t.desirability[4 * s + k] = t.desirability[4 * s + k] + *d; // :370-379
            }
        }
    }
    // ownership: |delta| agents sampled uniformly among the current non-owners (delta > 0) or owners (delta < 0)
    fn resample(current: &[u8], delta: i32) -> Vec<u8> {
        if delta == 0 {
            return Vec::new(); // unchanged
        }
        let want_owner = delta < 0;
        let pool: Vec<usize> = (0..current.len()).filter(|&i| (current[i] != 0) == want_owner).collect();
        let mut rng = thread_rng();
        let picked = sample_slice_ref(&mut rng, &pool, delta.abs() as usize); // panics if |delta| > pool, as :393-462 does
        let mut out = current.to_vec();
        for &&i in picked.iter() {
            out[i] = if want_owner { 0 } else { 1 };
        }
        out
    }
    Some(mvb::Intervention {
        day: iv.day,
        tables: Some(t),
        owns_bike: resample(&flat.owns_bike, iv.change_in_number_of_bikes),
        owns_car: resample(&flat.owns_car, iv.change_in_number_of_cars),
    })
}

/// Replaces `src/simulation.rs:94-136`: day-0 row, the day loop, one CSV row per weekday — same columns, same
/// order (`generate_csv_header`, `:207-211`).  `file` already holds the header.
pub fn run_on_gpu<W: Write>(id: &str, scenario: &Scenario, residents: &[Rc<RefCell<Agent>>], weather_pattern: &Vec<Weather>,
                            total_years: u32, gpu: i32, file: &mut W) -> Result<(), io::Error> {
    let flat = flatten(residents, scenario);
    let tables = tables_of(scenario);
    // uniform weights are never used here (flatten passes per-agent arrays); any finite values will do
    let prm = mvb::params(0.0, 0.0, 0.0, 1);
    let mut sim = mvb::Sim::create(&flat, &tables, &prm, gpu).unwrap_or_else(|e| panic!("[{}] {}", id, e));
    let iv = resolve_intervention(scenario, &flat);
    let weather: Vec<u8> = weather_pattern.iter().map(|w| (*w == Weather::Bad) as u8).collect();
    let n_days = total_years * 365; // loop bound of :101
    let ivs = [iv.as_ref()];
    let out = sim.run(&weather, n_days, &ivs).unwrap_or_else(|e| panic!("[{}] {}", id, e));
    let width = out.widths[0];
    for (k, day) in out.days.iter().enumerate() {
        let cols: Vec<String> = out.rows[0][k * width..(k + 1) * width].iter().map(|v| v.to_string()).collect();
        file.write(format!("{},{},{}\n", day, weather[*day as usize], cols.join(",")).as_bytes())?;
    }
    // write the final state back so that anything after the loop sees what the CPU path would have left
    let (habit, cur, last, norm) = sim.get_state(0, residents.len()).unwrap_or_else(|e| panic!("[{}] {}", id, e));
    for (i, cell) in residents.iter().enumerate() {
        let mut a = cell.borrow_mut();
        a.current_mode = MODES[cur[i] as usize];
        a.last_mode = MODES[last[i] as usize];
        a.norm = MODES[norm[i] as usize];
        for (k, m) in MODES.iter().enumerate() {
            if habit[4 * i + k] != 0.0 || a.habit.contains_key(m) {
                a.habit.insert(*m, habit[4 * i + k]);
            }
        }
    }
    Ok(())
}
