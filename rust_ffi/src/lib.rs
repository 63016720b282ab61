//! `extern "C"` declarations of include/ising_b200.h and a safe wrapper with the shape the
//! call sites in src/lattice.rs need.  Uncompiled here (no Rust toolchain in the build image).
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};

#[repr(C)]
pub struct ising_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct ising_graph {
    _p: [u8; 0],
}
#[repr(C)]
pub struct ising_sim {
    _p: [u8; 0],
}
#[repr(C)]
pub struct ising_comm {
    _private: [u8; 0],
}
#[repr(C)]
pub struct ising_pt {
    _p: [u8; 0],
}

pub const ISING_FLAG_PER_STEP_ENERGIES: u32 = 1 << 1;
pub const ISING_FLAG_LINEAR_SCHEDULE: u32 = 1 << 2;
pub const ISING_FLAG_EDGE_IMPORTANCE: u32 = 1 << 3;
pub const ISING_FLAG_NON_BASIC_MOVES: u32 = 1 << 4;

/// What one timestep consists of (ising_sim_set_moves): the nspinupdates / nedgeupdates /
/// nwormupdates of GraphState::do_time_step in units of whole passes (classicising.rs:100-106).
#[repr(C)]
pub struct ising_moves {
    pub struct_size: u32,
    pub spin_sweeps: u32,
    pub edge_passes: u32,
    pub worms: u32,
    pub worm_len: u32,
    pub edge_importance: u32,
}

#[repr(C)]
pub struct ising_run_args {
    pub struct_size: u32,
    pub flags: u32,
    pub beta: f64,
    pub sched_t: *const u64,
    pub sched_beta: *const f64,
    pub sched_len: u64,
    pub timesteps: u64,
    pub num_experiments: u64,
    pub thermalization: u64,
    pub sampling_freq: u64,
    pub seed: u64,
    pub replica_offset: u64,
    pub initial_state: *const u8,
}

extern "C" {
    pub fn ising_ctx_create(device: c_int, out: *mut *mut ising_ctx) -> c_int;
    pub fn ising_ctx_destroy(ctx: *mut ising_ctx);
    pub fn ising_last_error(ctx: *const ising_ctx) -> *const c_char;
    pub fn ising_graph_from_edges(
        ctx: *mut ising_ctx, nvars: u64, nedges: u64, a: *const u64, b: *const u64, j: *const f64,
        biases: *const f64, out: *mut *mut ising_graph,
    ) -> c_int;
    pub fn ising_graph_destroy(g: *mut ising_graph);
    pub fn ising_graph_get_edge_classes(g: *mut ising_graph, cls: *mut u32) -> c_int;
    pub fn ising_strong_edge_colouring(
        nvars: u64, nedges: u64, a: *const u64, b: *const u64, cls: *mut u32, nclasses: *mut u32,
    ) -> c_int;
    pub fn ising_make_seeds(seed_gen: u64, n: u64, out: *mut u64) -> c_int;
    pub fn ising_run_monte_carlo(
        ctx: *mut ising_ctx, g: *const ising_graph, args: *const ising_run_args, energies: *mut f64,
        states: *mut u8,
    ) -> c_int;
    pub fn ising_run_monte_carlo_sampling(
        ctx: *mut ising_ctx, g: *const ising_graph, args: *const ising_run_args, energies: *mut f64,
        states: *mut u8,
    ) -> c_int;
    pub fn ising_run_monte_carlo_annealing(
        ctx: *mut ising_ctx, g: *const ising_graph, args: *const ising_run_args, energies: *mut f64,
        states: *mut u8,
    ) -> c_int;
    // replay of a recorded (site, uniform) sequence: the bit-exact correctness mode
    pub fn ising_replay(
        ctx: *mut ising_ctx, g: *const ising_graph, beta: f64, num_experiments: u64, nattempts: u64,
        sites: *const u32, u: *const f64, init: *const u8, energies: *mut f64, states: *mut u8,
    ) -> c_int;

    // src/classicising.rs: device-resident experiments behind ClassicIsing
    pub fn ising_sim_create(
        ctx: *mut ising_ctx, g: *const ising_graph, num_experiments: u64, seed: u64, replica_offset: u64,
        out: *mut *mut ising_sim,
    ) -> c_int;
    pub fn ising_sim_destroy(sim: *mut ising_sim);
    pub fn ising_sim_set_moves(sim: *mut ising_sim, moves: *const ising_moves) -> c_int;
    pub fn ising_sim_step_acceptance(sim: *mut ising_sim, beta: f64, changed: *mut u64) -> c_int;
    pub fn ising_sim_set_states(sim: *mut ising_sim, states: *const u8) -> c_int;
    pub fn ising_sim_sweeps(
        sim: *mut ising_sim, betas: *const f64, nsweeps: u64, energies_per_sweep: *mut f64,
    ) -> c_int;
    pub fn ising_sim_run_sampling(
        sim: *mut ising_sim, beta: f64, thermalization: u64, sampling_freq: u64, n_samples: u64,
        energies: *mut f64, states: *mut u8,
    ) -> c_int;
    pub fn ising_sim_run_observables(
        sim: *mut ising_sim, beta: f64, thermalization: u64, sampling_freq: u64, n_samples: u64,
        energies: *mut f64, mags: *mut f64, overlaps: *mut f64,
    ) -> c_int;
    pub fn ising_sim_get_energies(sim: *mut ising_sim, energies: *mut f64) -> c_int;
    pub fn ising_sim_get_states(sim: *mut ising_sim, states: *mut u8) -> c_int;

    // src/tempering.rs: the replica loop of LatticeTempering for classical replicas
    pub fn ising_pt_create(
        ctx: *mut ising_ctx, g: *const ising_graph, betas: *const f64, nbetas: u64, cfg_lo: u64, cfg_hi: u64,
        seed: u64, out: *mut *mut ising_pt,
    ) -> c_int;
    pub fn ising_pt_destroy(pt: *mut ising_pt);
    pub fn ising_pt_sweeps(pt: *mut ising_pt, t: u64, local_energies: *mut f64) -> c_int;
    pub fn ising_pt_swap_step(pt: *mut ising_pt, all_energies: *const f64) -> c_int;
    pub fn ising_pt_timesteps_sample(
        pt: *mut ising_pt, timesteps: u64, replica_swap_freq: u64, sampling_freq: u64, states: *mut u8,
        energies: *mut f64,
    ) -> c_int;
    pub fn ising_pt_total_swaps(pt: *const ising_pt, out: *mut u64) -> c_int;
    pub fn ising_pt_get_pair_stats(pt: *const ising_pt, attempts: *mut u64, accepts: *mut u64) -> c_int;
    // multi-GPU: NCCL communicator owned by the library (one process per GPU); only the
    // ISING_COMM_ID_BYTES = 128 bytes of the unique id travel through the host program
    pub fn ising_comm_unique_id(out: *mut u8, capacity: u64) -> c_int;
    pub fn ising_comm_create(
        ctx: *mut ising_ctx, id: *const u8, rank: c_int, world: c_int, out: *mut *mut ising_comm,
    ) -> c_int;
    pub fn ising_comm_destroy(comm: *mut ising_comm);
    pub fn ising_comm_info(comm: *const ising_comm, rank: *mut c_int, world: *mut c_int) -> c_int;
    pub fn ising_pt_set_comm(pt: *mut ising_pt, comm: *mut ising_comm) -> c_int;
}

/// Owns a context + compiled graph; one per `Lattice` (rebuilt when biases change).
pub struct B200Graph {
    ctx: *mut ising_ctx,
    graph: *mut ising_graph,
}

impl B200Graph {
    pub fn new(device: i32, nvars: usize, edges: &[((usize, usize), f64)], biases: &[f64]) -> Result<Self, String> {
        let a: Vec<u64> = edges.iter().map(|((a, _), _)| *a as u64).collect();
        let b: Vec<u64> = edges.iter().map(|((_, b), _)| *b as u64).collect();
        let j: Vec<f64> = edges.iter().map(|(_, j)| *j).collect();
        unsafe {
            let mut ctx = std::ptr::null_mut();
            if ising_ctx_create(device, &mut ctx) != 0 {
                return Err(last_error(std::ptr::null()));
            }
            let mut graph = std::ptr::null_mut();
            let rc = ising_graph_from_edges(
                ctx, nvars as u64, edges.len() as u64, a.as_ptr(), b.as_ptr(), j.as_ptr(),
                if biases.iter().all(|b| *b == 0.0) { std::ptr::null() } else { biases.as_ptr() },
                &mut graph,
            );
            if rc != 0 {
                let msg = last_error(ctx);
                ising_ctx_destroy(ctx);
                return Err(msg);
            }
            Ok(B200Graph { ctx, graph })
        }
    }

    /// Fills `energies` [E] and `states` [E * nvars] (bool as u8), lattice.rs:171-221.
    pub fn run_monte_carlo(
        &self, beta: f64, timesteps: usize, seed: u64, initial_state: Option<&[bool]>, energies: &mut [f64],
        states: &mut [bool],
    ) -> Result<(), String> {
        let args = ising_run_args {
            struct_size: std::mem::size_of::<ising_run_args>() as u32,
            flags: 0,
            beta,
            sched_t: std::ptr::null(),
            sched_beta: std::ptr::null(),
            sched_len: 0,
            timesteps: timesteps as u64,
            num_experiments: energies.len() as u64,
            thermalization: 0,
            sampling_freq: 1,
            seed,
            replica_offset: 0,
            initial_state: initial_state.map_or(std::ptr::null(), |s| s.as_ptr() as *const u8),
        };
        let rc = unsafe {
            ising_run_monte_carlo(self.ctx, self.graph, &args, energies.as_mut_ptr(), states.as_mut_ptr() as *mut u8)
        };
        if rc == 0 { Ok(()) } else { Err(unsafe { last_error(self.ctx) }) }
    }
}

impl Drop for B200Graph {
    fn drop(&mut self) {
        unsafe {
            ising_graph_destroy(self.graph);
            ising_ctx_destroy(self.ctx);
        }
    }
}

unsafe fn last_error(ctx: *const ising_ctx) -> String {
    CStr::from_ptr(ising_last_error(ctx)).to_string_lossy().into_owned()
}
