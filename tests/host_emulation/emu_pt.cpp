// TEST INFRASTRUCTURE (CPU suite only).  Runs the kernels of the device-resident replica-exchange
// cycle (csrc/pt_device.cu: k_pt_cycle - energies from the bond counters, time average, swap
// decisions, slot maps, threshold tables - and k_pt_swap) from their own source on the host
// (cuda_on_host.h) for tests/test_device_source_on_host.py.  Together with the cluster sweep
// kernel (emu_stencil.cpp) that is a whole tempering run as ising_pt_timesteps_sample enqueues it,
// compared with oracle/msc_mirror.c (msc_mirror_pt).
#include "cuda_on_host.h"

#include <string.h>

#include "prepared/pt_device_kernels.cuh"

using namespace ising;

extern "C" void emu_pt_cycle(unsigned long long* nsat, double* e_local, double* e_all, uint32_t E, uint32_t e32,
                             double scale, unsigned long long nbonds, int mult, const uint32_t* gidx,
                             const double* betas, uint32_t* slot_of_cfg, uint32_t* cfg_of_slot, uint32_t R,
                             uint64_t seed, unsigned long long* stats, uint32_t* slot_of_replica, double* acc,
                             double t, int do_swap, const unsigned long long* t64, uint32_t W, int K,
                             uint32_t* tplane, uint32_t* tlow) {
    PtCycleArgs a;
    memset(&a, 0, sizeof a);
    a.nsat = nsat; a.e_local = e_local; a.e_all = e_all;
    a.E = E; a.e32 = e32; a.identity = 1u;
    a.scale = scale; a.nbonds = nbonds; a.mult = mult;
    a.gidx = gidx; a.betas = betas; a.slot_of_cfg = slot_of_cfg; a.cfg_of_slot = cfg_of_slot;
    a.R = R; a.key0 = (uint32_t)seed; a.key1 = (uint32_t)(seed >> 32);
    a.stats = stats; a.slot_of_replica = slot_of_replica; a.word_lo = 0;
    a.acc = acc; a.t = t; a.do_swap = do_swap;
    a.t64 = t64; a.W = W; a.K = K; a.tplane = tplane; a.tlow = tlow;
    emu::launch_v(k_pt_cycle, dim3(1), dim3(256), 0, a);
}

extern "C" void emu_pt_swap(const double* betas, const double* e_all, const uint32_t* gidx, uint32_t* slot_of_cfg,
                            uint32_t* cfg_of_slot, uint32_t R, uint64_t seed, unsigned long long* stats,
                            uint32_t* slot_of_replica, uint32_t e32) {
    emu::launch_v(k_pt_swap, dim3(1), dim3(256), 0, betas, e_all, gidx, slot_of_cfg, cfg_of_slot, R, (uint32_t)seed,
                  (uint32_t)(seed >> 32), stats, slot_of_replica, 0u, e32);
}
