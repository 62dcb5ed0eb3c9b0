//! dump_state.rs — drop-in for the reference crate (`src/dump_state.rs`): records what an UNMODIFIED CPU run of the
//! reference computed, so that motivate-b200's oracle can be pinned against the real thing
//! (`tests/test_reference_fixtures.py` consumes the result; `tools/make_ref_fixtures.sh` drives it).
//!
//! Enabled by the environment variable MOTIVATE_DUMP_DIR; nothing changes when it is unset.  All files are
//! little-endian, no padding.  Apply with the three one-line calls listed in README.md:
//!   simulation.rs after link_agents (:78)        dump_state::inputs(&id, &scenario, &residents, weather_pattern, total_years)
//!   simulation.rs after intervene (:104)         dump_state::intervention(&id, day, &scenario, &residents)
//!   simulation.rs after generate_csv_output (:134, and :98 with day 0)   dump_state::day(&id, day, &residents)
//!
//! DUMP/{id}/inputs.bin      "MVBD" u32 version=1, N, S, H, n_days, then: subculture u8[N], neighbourhood u16[N],
//!                           commute u8[N], owns_car u8[N], owns_bike u8[N], weather_sensitivity f32[N],
//!                           consistency, social_c, subculture_c, neighbourhood_c, average_weight f32[N] each,
//!                           habit f32[N][4], current_mode u8[N], norm u8[N],
//!                           social rowptr u64[N+1] + col u32[E], neighbour rowptr u64[N+1] + col u32[E],
//!                           supportiveness f32[H][4], capacity u32[H][4], desirability f32[S][4], weather u8[n_days]
//! DUMP/{id}/intervention.bin  day u32, supportiveness f32[H][4], capacity u32[H][4], desirability f32[S][4],
//!                           owns_car u8[N], owns_bike u8[N]            (the state right after `intervene`)
//! DUMP/{id}/days.bin        per emitted row: day u32, current_mode u8[N], norm u8[N], habit f32[N][4]

use std::cell::RefCell;
use std::collections::HashMap;
use std::env;
use std::fs;
use std::fs::OpenOptions;
use std::io::Write;
use std::rc::Rc;

use agent::Agent;
use journey_type::JourneyType;
use scenario::Scenario;
use transport_mode::TransportMode;
use weather::Weather;

const MODES: [TransportMode; 4] = [TransportMode::Car, TransportMode::PublicTransport, TransportMode::Cycle, TransportMode::Walk];

fn dir(id: &str) -> Option<String> {
    env::var("MOTIVATE_DUMP_DIR").ok().map(|d| {
        let p = format!("{}/{}", d, id);
        fs::create_dir_all(&p).expect("cannot create the dump directory");
        p
    })
}
fn code(m: TransportMode) -> u8 { MODES.iter().position(|x| *x == m).unwrap() as u8 }
fn u32le(v: u32) -> [u8; 4] { [v as u8, (v >> 8) as u8, (v >> 16) as u8, (v >> 24) as u8] }
fn put_u32(b: &mut Vec<u8>, v: u32) { b.extend_from_slice(&u32le(v)); }
fn put_u64(b: &mut Vec<u8>, v: u64) { put_u32(b, v as u32); put_u32(b, (v >> 32) as u32); }
fn put_f32(b: &mut Vec<u8>, v: f32) { put_u32(b, v.to_bits()); }

fn tables(b: &mut Vec<u8>, scenario: &Scenario) {
    for nb in scenario.neighbourhoods.iter() { for m in MODES.iter() { put_f32(b, *nb.supportiveness.borrow().get(m).unwrap()); } }
    for nb in scenario.neighbourhoods.iter() { for m in MODES.iter() { put_u32(b, *nb.capacity.borrow().get(m).unwrap()); } }
    for sc in scenario.subcultures.iter() { for m in MODES.iter() { put_f32(b, *sc.desirability.borrow().get(m).unwrap()); } }
}

pub fn inputs(id: &str, scenario: &Scenario, residents: &[Rc<RefCell<Agent>>], weather_pattern: &Vec<Weather>, total_years: u32) {
    let d = match dir(id) { Some(d) => d, None => return };
    let n = residents.len();
    let index_of: HashMap<*const RefCell<Agent>, u32> =
        residents.iter().enumerate().map(|(i, a)| (&**a as *const RefCell<Agent>, i as u32)).collect();
    let mut b: Vec<u8> = Vec::new();
    b.extend_from_slice(b"MVBD");
    put_u32(&mut b, 1);
    put_u32(&mut b, n as u32);
    put_u32(&mut b, scenario.subcultures.len() as u32);
    put_u32(&mut b, scenario.neighbourhoods.len() as u32);
    put_u32(&mut b, total_years * 365);
    for a in residents { b.push(scenario.subcultures.iter().position(|s| s.id == a.borrow().subculture.id).unwrap() as u8); }
    for a in residents {
        let h = scenario.neighbourhoods.iter().position(|s| s.id == a.borrow().neighbourhood.id).unwrap() as u16;
        b.push(h as u8); b.push((h >> 8) as u8);
    }
    for a in residents {
        b.push(match a.borrow().commute_length { JourneyType::LocalCommute => 0, JourneyType::CityCommute => 1, JourneyType::DistantCommute => 2 });
    }
    for a in residents { b.push(a.borrow().owns_car as u8); }
    for a in residents { b.push(a.borrow().owns_bike as u8); }
    for a in residents { put_f32(&mut b, a.borrow().weather_sensitivity); }
    for a in residents { put_f32(&mut b, a.borrow().consistency); }
    for a in residents { put_f32(&mut b, a.borrow().social_connectivity); }
    for a in residents { put_f32(&mut b, a.borrow().subculture_connectivity); }
    for a in residents { put_f32(&mut b, a.borrow().neighbourhood_connectivity); }
    for a in residents { put_f32(&mut b, a.borrow().average_weight); }
    for a in residents { for m in MODES.iter() { put_f32(&mut b, *a.borrow().habit.get(m).unwrap_or(&0.0)); } }
    for a in residents { b.push(code(a.borrow().current_mode)); }
    for a in residents { b.push(code(a.borrow().norm)); }
    for pass in 0..2 {
        let mut at: u64 = 0;
        put_u64(&mut b, 0);
        for a in residents { at += if pass == 0 { a.borrow().social_network.len() } else { a.borrow().neighbours.len() } as u64; put_u64(&mut b, at); }
        for a in residents {
            let a = a.borrow();
            for f in (if pass == 0 { &a.social_network } else { &a.neighbours }).iter() { put_u32(&mut b, index_of[&(&**f as *const RefCell<Agent>)]); }
        }
    }
    tables(&mut b, scenario);
    for w in weather_pattern.iter().take((total_years * 365) as usize) { b.push((*w == Weather::Bad) as u8); }
    fs::File::create(format!("{}/inputs.bin", d)).unwrap().write_all(&b).unwrap();
    let _ = fs::remove_file(format!("{}/days.bin", d));
}

pub fn intervention(id: &str, day: u32, scenario: &Scenario, residents: &[Rc<RefCell<Agent>>]) {
    let d = match dir(id) { Some(d) => d, None => return };
    let mut b: Vec<u8> = Vec::new();
    put_u32(&mut b, day);
    tables(&mut b, scenario);
    for a in residents { b.push(a.borrow().owns_car as u8); }
    for a in residents { b.push(a.borrow().owns_bike as u8); }
    fs::File::create(format!("{}/intervention.bin", d)).unwrap().write_all(&b).unwrap();
}

pub fn day(id: &str, day: u32, residents: &[Rc<RefCell<Agent>>]) {
    let d = match dir(id) { Some(d) => d, None => return };
    let mut b: Vec<u8> = Vec::with_capacity(4 + residents.len() * 18);
    put_u32(&mut b, day);
    for a in residents { b.push(code(a.borrow().current_mode)); }
    for a in residents { b.push(code(a.borrow().norm)); }
    for a in residents { for m in MODES.iter() { put_f32(&mut b, *a.borrow().habit.get(m).unwrap_or(&0.0)); } }
    OpenOptions::new().create(true).append(true).open(format!("{}/days.bin", d)).unwrap().write_all(&b).unwrap();
}
