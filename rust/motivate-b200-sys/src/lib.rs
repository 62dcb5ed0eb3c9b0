//! motivate-b200-sys — Rust side of the C ABI in `include/motivate_b200.h` (ABI version 1).
//!
//! The reference (0xr0bert/Motivate) has no FFI; this crate is what its `simulation::run`
//! (`src/simulation.rs:40-149`) binds in order to hand the weekday loop (`:101-136`) to the B200 library.
//! `raw` declares every entry point of the header, field for field and argument for argument
//! (`tests/test_rust_binding.py` of the motivate-b200 repo parses this file and the header and compares
//! them); the rest of the file is a small safe wrapper: owned structure-of-arrays population, tables,
//! a `Sim` handle that is destroyed on drop, errors as `Result<_, Error>`.
//!
//! NOTE: this image has no cargo/rustc, so the crate is delivered as source and has not been compiled
//! here; the same calls are exercised through ctypes (`motivate_b200/_ffi.py`), a C99 consumer
//! (`tests/c/abi_consumer.c`) and the C++ host (`motivate_b200/csrc/host`).
#![allow(non_camel_case_types)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};
use std::ptr;

pub const MVB_ABI_VERSION: c_int = 1;
pub const MVB_OK: c_int = 0;
pub const MVB_ERR_ARG: c_int = -1;
pub const MVB_ERR_CUDA: c_int = -2;
pub const MVB_ERR_NOMEM: c_int = -3;
pub const MVB_ERR_STATE: c_int = -4;
pub const MVB_ERR_DOMAIN: c_int = -5;
pub const MVB_ERR_COMM: c_int = -6;
pub const MVB_N_MODES: usize = 4;
pub const MVB_N_COMMUTES: usize = 3;
pub const MVB_STATS_FIXED: usize = 7;
pub const MVB_UNIQUE_ID_BYTES: usize = 128;
pub const MVB_FLAG_NONE: u32 = 0;
pub const MVB_EXCHANGE_NONE: c_int = 0;
pub const MVB_EXCHANGE_NCCL: c_int = 1;
pub const MVB_EXCHANGE_P2P: c_int = 2;

/// Everything `extern "C"`: names, argument order and struct field order follow the header exactly.
pub mod raw {
    use std::os::raw::{c_char, c_int, c_void};

    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct mvb_population {
        pub n_agents: u32,
        pub subculture: *const u8,
        pub neighbourhood: *const u16,
        pub commute: *const u8,
        pub owns_car: *const u8,
        pub owns_bike: *const u8,
        pub weather_sensitivity: *const f32,
        pub consistency: *const f32,
        pub social_connectivity: *const f32,
        pub subculture_connectivity: *const f32,
        pub neighbourhood_connectivity: *const f32,
        pub average_weight: *const f32,
        pub habit: *const f32,
        pub current_mode: *const u8,
        pub norm: *const u8,
        pub social_rowptr: *const u64,
        pub social_col: *const u32,
        pub neighbour_rowptr: *const u64,
        pub neighbour_col: *const u32,
    }

    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct mvb_tables {
        pub n_subcultures: u32,
        pub n_neighbourhoods: u32,
        pub supportiveness: *const f32,
        pub capacity: *const u32,
        pub desirability: *const f32,
    }

    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct mvb_params {
        pub consistency: f32,
        pub social_connectivity: f32,
        pub subculture_connectivity: f32,
        pub neighbourhood_connectivity: f32,
        pub average_weight: f32,
        pub flags: u32,
    }

    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct mvb_intervention {
        pub day: u32,
        pub tables: *const mvb_tables,
        pub owns_car: *const u8,
        pub owns_bike: *const u8,
    }

    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct mvb_synth_spec {
        pub n_agents: u32,
        pub n_subcultures: u32,
        pub n_neighbourhoods: u32,
        pub social_links_min: u32,
        pub neighbour_links_min: u32,
        pub n_cars: u32,
        pub n_bikes: u32,
        pub seed: u64,
        pub n_gmm: u32,
        pub gmm: [[f64; 3]; 8],
    }

    /// Opaque handles.
    pub enum mvb_sim {}
    pub enum mvb_synth {}

    extern "C" {
        pub fn mvb_abi_version() -> c_int;
        pub fn mvb_device_count() -> c_int;
        pub fn mvb_last_error() -> *const c_char;
        pub fn mvb_create(pop: *const mvb_population, tab: *const mvb_tables, prm: *const mvb_params, device: c_int, out: *mut *mut mvb_sim) -> c_int;
        pub fn mvb_create_batch(n_replicas: u32, pops: *const *const mvb_population, tabs: *const *const mvb_tables, prm: *const mvb_params, device: c_int, out: *mut *mut mvb_sim) -> c_int;
        pub fn mvb_destroy(sim: *mut mvb_sim);
        pub fn mvb_n_replicas(sim: *const mvb_sim) -> u32;
        pub fn mvb_stats_width(sim: *const mvb_sim, replica: u32) -> u32;
        pub fn mvb_step(sim: *mut mvb_sim, weather_today: u8, weather_changed: u8) -> c_int;
        pub fn mvb_run(sim: *mut mvb_sim, weather: *const u8, n_days: u32, ivs: *const *const mvb_intervention, days_out: *mut u32, rows_out: *mut u32) -> c_int;
        pub fn mvb_run_rows(n_days: u32) -> u32;
        pub fn mvb_run_async(sim: *mut mvb_sim, weather: *const u8, first_day: u32, end_day: u32, ivs: *const *const mvb_intervention) -> c_int;
        pub fn mvb_fetch_rows(sim: *mut mvb_sim, n_rows: *mut u32, days_out: *mut u32, rows_out: *mut u32, rows_capacity_u32: usize) -> c_int;
        pub fn mvb_reset_rows(sim: *mut mvb_sim) -> c_int;
        pub fn mvb_sync(sim: *mut mvb_sim) -> c_int;
        pub fn mvb_stream(sim: *mut mvb_sim) -> *mut c_void;
        pub fn mvb_last_run_stats(sim: *mut mvb_sim, device_ms: *mut f32, kernel_launches: *mut u64) -> c_int;
        pub fn mvb_set_tables(sim: *mut mvb_sim, replica: u32, tab: *const mvb_tables) -> c_int;
        pub fn mvb_set_ownership(sim: *mut mvb_sim, replica: u32, owns_car: *const u8, owns_bike: *const u8) -> c_int;
        pub fn mvb_stats_row(sim: *mut mvb_sim, replica: u32, row_out: *mut u32) -> c_int;
        pub fn mvb_get_state(sim: *mut mvb_sim, replica: u32, habit: *mut f32, current_mode: *mut u8, last_mode: *mut u8, norm: *mut u8, hist: *mut u32, congestion: *mut f32) -> c_int;
        pub fn mvb_algorithmic_bytes_per_day(sim: *const mvb_sim, replica: u32) -> u64;
        pub fn mvb_comm_unique_id(id_out: *mut c_void) -> c_int;
        pub fn mvb_create_partitioned(pop: *const mvb_population, tab: *const mvb_tables, prm: *const mvb_params, device: c_int, rank: u32, world: u32, unique_id: *const c_void, out: *mut *mut mvb_sim) -> c_int;
        pub fn mvb_partition_owner(sim: *const mvb_sim, rank_out: *mut u32) -> c_int;
        pub fn mvb_partition_plan(pop: *const mvb_population, n_subcultures: u32, n_neighbourhoods: u32, world: u32, rank_out: *mut u32) -> c_int;
        pub fn mvb_exchange_mode(sim: *const mvb_sim) -> c_int;
        pub fn mvb_synth_create(spec: *const mvb_synth_spec, out: *mut *mut mvb_synth) -> c_int;
        pub fn mvb_synth_create_device(spec: *const mvb_synth_spec, device: c_int, out: *mut *mut mvb_synth) -> c_int;
        pub fn mvb_synth_copy_to_host(device_population: *const mvb_synth, host_out: *mut *mut mvb_synth) -> c_int;
        pub fn mvb_synth_destroy(s: *mut mvb_synth);
        pub fn mvb_synth_population(s: *const mvb_synth) -> *const mvb_population;
        pub fn mvb_synth_ba_network(min_links: u32, n_nodes: u32, seed: u64, rowptr_out: *mut *mut u64, col_out: *mut *mut u32) -> c_int;
        pub fn mvb_free(p: *mut c_void);
    }
}

// ---------------------------------------------------------------------------------------------------
// Safe wrapper
// ---------------------------------------------------------------------------------------------------

/// A non-zero return code of the library with the text of `mvb_last_error()`.  These replace the
/// reference's panics (`expect` at `src/simulation.rs:318,368`, `unwrap` at `:232-234`).
#[derive(Debug, Clone)]
pub struct Error {
    pub code: c_int,
    pub message: String,
}

impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter) -> std::fmt::Result {
        write!(f, "motivate_b200 error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for Error {
    fn description(&self) -> &str {
        &self.message
    }
}

fn check(rc: c_int) -> Result<(), Error> {
    if rc == MVB_OK {
        return Ok(());
    }
    let p: *const c_char = unsafe { raw::mvb_last_error() };
    let message = if p.is_null() { String::new() } else { unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned() };
    Err(Error { code: rc, message: message })
}

/// Mode / commute / weather codes (`src/transport_mode.rs:3-8`, `src/journey_type.rs:5-9`, `src/weather.rs:7-10`).
pub mod codes {
    pub const CAR: u8 = 0;
    pub const PUBLIC_TRANSPORT: u8 = 1;
    pub const CYCLE: u8 = 2;
    pub const WALK: u8 = 3;
    pub const LOCAL_COMMUTE: u8 = 0;
    pub const CITY_COMMUTE: u8 = 1;
    pub const DISTANT_COMMUTE: u8 = 2;
    pub const GOOD: u8 = 0;
    pub const BAD: u8 = 1;
}

/// One population as owned structure-of-arrays (what `load_unlinked_agents_from_file` + `link_agents`
/// leave in memory, `src/agent_generation.rs:26-55`, `src/simulation.rs:155-189`).  Agent ids are positions.
#[derive(Default, Clone)]
pub struct Population {
    pub subculture: Vec<u8>,
    pub neighbourhood: Vec<u16>,
    pub commute: Vec<u8>,
    pub owns_car: Vec<u8>,
    pub owns_bike: Vec<u8>,
    pub weather_sensitivity: Vec<f32>,
    /// Per-agent weights; an empty vector means "uniform, take it from `Params`".
    pub consistency: Vec<f32>,
    pub social_connectivity: Vec<f32>,
    pub subculture_connectivity: Vec<f32>,
    pub neighbourhood_connectivity: Vec<f32>,
    pub average_weight: Vec<f32>,
    /// `[N][4]` dense; an absent `HashMap` key is 0.0 (`src/agent.rs:62`).
    pub habit: Vec<f32>,
    pub current_mode: Vec<u8>,
    pub norm: Vec<u8>,
    /// CSR of `agent.social_network` / `agent.neighbours`, duplicates kept (`src/agent.rs:322-332`).
    pub social_rowptr: Vec<u64>,
    pub social_col: Vec<u32>,
    pub neighbour_rowptr: Vec<u64>,
    pub neighbour_col: Vec<u32>,
}

fn opt<T>(v: &[T]) -> *const T {
    if v.is_empty() { ptr::null() } else { v.as_ptr() }
}

impl Population {
    pub fn n_agents(&self) -> usize {
        self.subculture.len()
    }
    /// Borrowed view for the duration of a call (the library copies what it needs).
    pub fn as_raw(&self) -> raw::mvb_population {
        raw::mvb_population {
            n_agents: self.n_agents() as u32,
            subculture: self.subculture.as_ptr(),
            neighbourhood: self.neighbourhood.as_ptr(),
            commute: self.commute.as_ptr(),
            owns_car: self.owns_car.as_ptr(),
            owns_bike: self.owns_bike.as_ptr(),
            weather_sensitivity: self.weather_sensitivity.as_ptr(),
            consistency: opt(&self.consistency),
            social_connectivity: opt(&self.social_connectivity),
            subculture_connectivity: opt(&self.subculture_connectivity),
            neighbourhood_connectivity: opt(&self.neighbourhood_connectivity),
            average_weight: opt(&self.average_weight),
            habit: self.habit.as_ptr(),
            current_mode: self.current_mode.as_ptr(),
            norm: self.norm.as_ptr(),
            social_rowptr: self.social_rowptr.as_ptr(),
            social_col: opt(&self.social_col),
            neighbour_rowptr: self.neighbour_rowptr.as_ptr(),
            neighbour_col: opt(&self.neighbour_col),
        }
    }
}

/// Scenario tables in `TransportMode` order (`src/neighbourhood.rs:13-31`, `src/subculture.rs:9-14`).
#[derive(Default, Clone)]
pub struct Tables {
    pub n_subcultures: u32,
    pub n_neighbourhoods: u32,
    pub supportiveness: Vec<f32>, // [H][4]
    pub capacity: Vec<u32>,       // [H][4]
    pub desirability: Vec<f32>,   // [S][4]
}

impl Tables {
    pub fn as_raw(&self) -> raw::mvb_tables {
        raw::mvb_tables {
            n_subcultures: self.n_subcultures,
            n_neighbourhoods: self.n_neighbourhoods,
            supportiveness: self.supportiveness.as_ptr(),
            capacity: self.capacity.as_ptr(),
            desirability: self.desirability.as_ptr(),
        }
    }
}

/// Uniform weights (`src/main.rs:194-198`, `src/agent_generation.rs:226,246`).
pub fn params(social: f32, subculture: f32, neighbourhood: f32, days_in_habit_average: u32) -> raw::mvb_params {
    raw::mvb_params {
        consistency: 1.0,
        social_connectivity: social,
        subculture_connectivity: subculture,
        neighbourhood_connectivity: neighbourhood,
        average_weight: 2.0f32 / (days_in_habit_average as f32 + 1.0f32),
        flags: MVB_FLAG_NONE,
    }
}

/// The effect of `intervene` (`src/simulation.rs:304-464`) from `day` on, resolved by the host.
#[derive(Default, Clone)]
pub struct Intervention {
    pub day: u32,
    pub tables: Option<Tables>,
    pub owns_car: Vec<u8>,  // empty = unchanged
    pub owns_bike: Vec<u8>, // empty = unchanged
}

/// Statistics rows of one run: `days[k]` is the Day column of row `k`; `rows[r]` holds replica r's rows,
/// `width(r)` u32 each, in the column order of `src/simulation.rs:207-211` without Day and Rain.
pub struct Rows {
    pub days: Vec<u32>,
    pub rows: Vec<Vec<u32>>,
    pub widths: Vec<usize>,
}

/// R replicas resident on one GPU and stepped together.  Single-threaded; distinct handles may live on
/// distinct threads (one rayon worker per GPU instead of per replica, `src/main.rs:89-121`).
pub struct Sim {
    h: *mut raw::mvb_sim,
}

unsafe impl Send for Sim {}

impl Sim {
    pub fn abi_version() -> i32 {
        unsafe { raw::mvb_abi_version() }
    }
    pub fn device_count() -> i32 {
        unsafe { raw::mvb_device_count() }
    }

    pub fn create(pop: &Population, tab: &Tables, prm: &raw::mvb_params, device: i32) -> Result<Sim, Error> {
        let (p, t) = (pop.as_raw(), tab.as_raw());
        let mut h = ptr::null_mut();
        check(unsafe { raw::mvb_create(&p, &t, prm, device, &mut h) })?;
        Ok(Sim { h: h })
    }

    /// All replicas of one GPU in one call.  Populations that are the same object (pointer equality of the
    /// graph arrays) share one copy of the link data on the device (scenario sweeps).
    pub fn create_batch(pops: &[&Population], tabs: &[&Tables], prm: &raw::mvb_params, device: i32) -> Result<Sim, Error> {
        assert_eq!(pops.len(), tabs.len());
        let praw: Vec<raw::mvb_population> = pops.iter().map(|p| p.as_raw()).collect();
        let traw: Vec<raw::mvb_tables> = tabs.iter().map(|t| t.as_raw()).collect();
        let pp: Vec<*const raw::mvb_population> = praw.iter().map(|p| p as *const _).collect();
        let tp: Vec<*const raw::mvb_tables> = traw.iter().map(|t| t as *const _).collect();
        let mut h = ptr::null_mut();
        check(unsafe { raw::mvb_create_batch(pops.len() as u32, pp.as_ptr(), tp.as_ptr(), prm, device, &mut h) })?;
        Ok(Sim { h: h })
    }

    /// One population over `world` GPUs, one process per GPU (BASELINE config 4).
    pub fn unique_id() -> Result<[u8; MVB_UNIQUE_ID_BYTES], Error> {
        let mut id = [0u8; MVB_UNIQUE_ID_BYTES];
        check(unsafe { raw::mvb_comm_unique_id(id.as_mut_ptr() as *mut c_void) })?;
        Ok(id)
    }
    pub fn create_partitioned(pop: &Population, tab: &Tables, prm: &raw::mvb_params, device: i32, rank: u32, world: u32,
                              unique_id: &[u8; MVB_UNIQUE_ID_BYTES]) -> Result<Sim, Error> {
        let (p, t) = (pop.as_raw(), tab.as_raw());
        let mut h = ptr::null_mut();
        check(unsafe { raw::mvb_create_partitioned(&p, &t, prm, device, rank, world, unique_id.as_ptr() as *const c_void, &mut h) })?;
        Ok(Sim { h: h })
    }
    pub fn partition_owner(&self, n_agents: usize) -> Result<Vec<u32>, Error> {
        let mut out = vec![0u32; n_agents];
        check(unsafe { raw::mvb_partition_owner(self.h, out.as_mut_ptr()) })?;
        Ok(out)
    }
    pub fn partition_plan(pop: &Population, n_subcultures: u32, n_neighbourhoods: u32, world: u32) -> Result<Vec<u32>, Error> {
        let p = pop.as_raw();
        let mut out = vec![0u32; pop.n_agents()];
        check(unsafe { raw::mvb_partition_plan(&p, n_subcultures, n_neighbourhoods, world, out.as_mut_ptr()) })?;
        Ok(out)
    }

    /// How the ranks of an agent-partitioned handle exchange modes and counters: MVB_EXCHANGE_P2P (fused into the
    /// weekday kernel, over peer memory), MVB_EXCHANGE_NCCL (all-gather + all-reduce) or MVB_EXCHANGE_NONE.
    pub fn exchange_mode(&self) -> c_int {
        unsafe { raw::mvb_exchange_mode(self.h) }
    }

    pub fn n_replicas(&self) -> usize {
        unsafe { raw::mvb_n_replicas(self.h) as usize }
    }
    pub fn stats_width(&self, replica: usize) -> usize {
        unsafe { raw::mvb_stats_width(self.h, replica as u32) as usize }
    }
    pub fn run_rows(n_days: u32) -> usize {
        unsafe { raw::mvb_run_rows(n_days) as usize }
    }

    /// The loop of `src/simulation.rs:95-136` for every replica: day-0 row, then for `day in 1..n_days`
    /// the intervention (if it is its day), and on weekdays one step + one row.
    pub fn run(&mut self, weather: &[u8], n_days: u32, ivs: &[Option<&Intervention>]) -> Result<Rows, Error> {
        assert!(weather.len() >= n_days as usize);
        let r = self.n_replicas();
        assert!(ivs.is_empty() || ivs.len() == r);
        // keep the raw views alive until the call returns
        let tabs: Vec<Option<raw::mvb_tables>> = ivs.iter().map(|iv| iv.and_then(|i| i.tables.as_ref().map(|t| t.as_raw()))).collect();
        let raws: Vec<Option<raw::mvb_intervention>> = ivs.iter().zip(tabs.iter()).map(|(iv, t)| {
            iv.map(|i| raw::mvb_intervention {
                day: i.day,
                tables: t.as_ref().map(|x| x as *const _).unwrap_or(ptr::null()),
                owns_car: opt(&i.owns_car),
                owns_bike: opt(&i.owns_bike),
            })
        }).collect();
        let ptrs: Vec<*const raw::mvb_intervention> = raws.iter().map(|o| o.as_ref().map(|x| x as *const _).unwrap_or(ptr::null())).collect();
        let n_rows = Sim::run_rows(n_days);
        let widths: Vec<usize> = (0..r).map(|k| self.stats_width(k)).collect();
        let total: usize = widths.iter().map(|w| w * n_rows).sum();
        let mut days = vec![0u32; n_rows];
        let mut flat = vec![0u32; total];
        check(unsafe {
            raw::mvb_run(self.h, weather.as_ptr(), n_days, if ptrs.is_empty() { ptr::null() } else { ptrs.as_ptr() },
                         days.as_mut_ptr(), flat.as_mut_ptr())
        })?;
        let mut rows = Vec::with_capacity(r);
        let mut at = 0;
        for w in &widths {
            rows.push(flat[at..at + w * n_rows].to_vec());
            at += w * n_rows;
        }
        Ok(Rows { days: days, rows: rows, widths: widths })
    }

    /// One weekday (`src/simulation.rs:116-128`); asynchronous on the handle's stream.
    pub fn step(&mut self, weather_today: u8, weather_changed: bool) -> Result<(), Error> {
        check(unsafe { raw::mvb_step(self.h, weather_today, weather_changed as u8) })
    }
    /// Days `first_day..end_day` of the calendar without copying anything back.
    pub fn run_async(&mut self, weather: &[u8], first_day: u32, end_day: u32, ivs: &[*const raw::mvb_intervention]) -> Result<(), Error> {
        check(unsafe { raw::mvb_run_async(self.h, weather.as_ptr(), first_day, end_day, if ivs.is_empty() { ptr::null() } else { ivs.as_ptr() }) })
    }
    pub fn fetch_rows(&mut self) -> Result<Rows, Error> {
        let mut n: u32 = 0;
        check(unsafe { raw::mvb_fetch_rows(self.h, &mut n, ptr::null_mut(), ptr::null_mut(), 0) })?;
        let r = self.n_replicas();
        let widths: Vec<usize> = (0..r).map(|k| self.stats_width(k)).collect();
        let total: usize = widths.iter().map(|w| w * n as usize).sum();
        let mut days = vec![0u32; n as usize];
        let mut flat = vec![0u32; total];
        check(unsafe { raw::mvb_fetch_rows(self.h, &mut n, days.as_mut_ptr(), flat.as_mut_ptr(), total) })?;
        let mut rows = Vec::with_capacity(r);
        let mut at = 0;
        for w in &widths {
            rows.push(flat[at..at + w * n as usize].to_vec());
            at += w * n as usize;
        }
        Ok(Rows { days: days, rows: rows, widths: widths })
    }
    pub fn reset_rows(&mut self) -> Result<(), Error> {
        check(unsafe { raw::mvb_reset_rows(self.h) })
    }
    pub fn sync(&mut self) -> Result<(), Error> {
        check(unsafe { raw::mvb_sync(self.h) })
    }
    /// `cudaStream_t` of the handle.
    pub fn stream(&mut self) -> *mut c_void {
        unsafe { raw::mvb_stream(self.h) }
    }
    /// (device milliseconds, day-kernel launches) of the last run.
    pub fn last_run_stats(&mut self) -> Result<(f32, u64), Error> {
        let (mut ms, mut n) = (0f32, 0u64);
        check(unsafe { raw::mvb_last_run_stats(self.h, &mut ms, &mut n) })?;
        Ok((ms, n))
    }

    pub fn set_tables(&mut self, replica: usize, tab: &Tables) -> Result<(), Error> {
        let t = tab.as_raw();
        check(unsafe { raw::mvb_set_tables(self.h, replica as u32, &t) })
    }
    pub fn set_ownership(&mut self, replica: usize, owns_car: Option<&[u8]>, owns_bike: Option<&[u8]>) -> Result<(), Error> {
        check(unsafe {
            raw::mvb_set_ownership(self.h, replica as u32, owns_car.map(|v| v.as_ptr()).unwrap_or(ptr::null()),
                                   owns_bike.map(|v| v.as_ptr()).unwrap_or(ptr::null()))
        })
    }
    pub fn stats_row(&mut self, replica: usize) -> Result<Vec<u32>, Error> {
        let mut row = vec![0u32; self.stats_width(replica)];
        check(unsafe { raw::mvb_stats_row(self.h, replica as u32, row.as_mut_ptr()) })?;
        Ok(row)
    }
    /// (habit `[N][4]`, current_mode, last_mode, norm) in the caller's agent order.
    pub fn get_state(&mut self, replica: usize, n_agents: usize) -> Result<(Vec<f32>, Vec<u8>, Vec<u8>, Vec<u8>), Error> {
        let mut habit = vec![0f32; 4 * n_agents];
        let (mut cur, mut last, mut norm) = (vec![0u8; n_agents], vec![0u8; n_agents], vec![0u8; n_agents]);
        check(unsafe {
            raw::mvb_get_state(self.h, replica as u32, habit.as_mut_ptr(), cur.as_mut_ptr(), last.as_mut_ptr(), norm.as_mut_ptr(),
                               ptr::null_mut(), ptr::null_mut())
        })?;
        Ok((habit, cur, last, norm))
    }
    pub fn algorithmic_bytes_per_day(&self, replica: usize) -> u64 {
        unsafe { raw::mvb_algorithmic_bytes_per_day(self.h, replica as u32) }
    }
}

impl Drop for Sim {
    fn drop(&mut self) {
        unsafe { raw::mvb_destroy(self.h) }
    }
}

/// Seeded synthetic population following `src/agent_generation.rs:114-325` / `src/social_network.rs:12-48`.
pub struct Synth {
    h: *mut raw::mvb_synth,
}

impl Synth {
    pub fn create(spec: &raw::mvb_synth_spec) -> Result<Synth, Error> {
        let mut h = ptr::null_mut();
        check(unsafe { raw::mvb_synth_create(spec, &mut h) })?;
        Ok(Synth { h: h })
    }
    /// Generated on GPU `device`; `population()` then holds DEVICE pointers, which `mvb_create*` accepts as they are.
    pub fn create_on_device(spec: &raw::mvb_synth_spec, device: i32) -> Result<Synth, Error> {
        let mut h = ptr::null_mut();
        check(unsafe { raw::mvb_synth_create_device(spec, device, &mut h) })?;
        Ok(Synth { h: h })
    }
    pub fn copy_to_host(&self) -> Result<Synth, Error> {
        let mut h = ptr::null_mut();
        check(unsafe { raw::mvb_synth_copy_to_host(self.h, &mut h) })?;
        Ok(Synth { h: h })
    }
    /// Borrowed view; valid while `self` lives.
    pub fn population(&self) -> &raw::mvb_population {
        unsafe { &*raw::mvb_synth_population(self.h) }
    }
}
impl Drop for Synth {
    fn drop(&mut self) {
        unsafe { raw::mvb_synth_destroy(self.h) }
    }
}

/// Barabási–Albert network as `src/social_network.rs:12-48` builds it, as CSR `(rowptr, col)`.
pub fn ba_network(min_links: u32, n_nodes: u32, seed: u64) -> Result<(Vec<u64>, Vec<u32>), Error> {
    let (mut rp, mut cl): (*mut u64, *mut u32) = (ptr::null_mut(), ptr::null_mut());
    check(unsafe { raw::mvb_synth_ba_network(min_links, n_nodes, seed, &mut rp, &mut cl) })?;
    let rowptr = unsafe { std::slice::from_raw_parts(rp, n_nodes as usize + 1) }.to_vec();
    let col = unsafe { std::slice::from_raw_parts(cl, rowptr[n_nodes as usize] as usize) }.to_vec();
    unsafe {
        raw::mvb_free(rp as *mut c_void);
        raw::mvb_free(cl as *mut c_void);
    }
    Ok((rowptr, col))
}
