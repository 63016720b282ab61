// C ABI: per-experiment inverse temperatures and classical parallel tempering (ising_pt_*).
#include "api_internal.h"

// ------------------------------------------------------------------------------------------
// per-experiment inverse temperatures and classical parallel tempering
// ------------------------------------------------------------------------------------------
extern "C" int ising_sim_set_betas(ising_sim* s, const double* betas) {
    CtxLock _lk(s ? s->ctx : nullptr);
    if (!s || !betas) return fail(s ? s->ctx : nullptr, ISING_E_INVALID, "sim/betas is NULL");
    ising_ctx* ctx = s->ctx;
    if (s->real)
        return fail(ctx, ISING_E_UNSUPPORTED,
                    "per-experiment betas need integer energy classes (all |J| equal, no bias)");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const HostGraph& h = s->g->h;
    const uint32_t W = s->lay.W, E32 = 32 * W;
    if (!s->general) {
        // checkerboard layout: classes dE = 4|J|, 8|J| (, 12|J|); tables are built for 6 planes
        if (s->planes != 6)
            return fail(ctx, ISING_E_UNSUPPORTED, "per-experiment betas on a lattice need planes = 6");
        std::vector<unsigned long long> t64((size_t)E32 * 3, 0ull);
        const int ncls = h.kind == ISING_KIND_STENCIL3D ? 3 : 2;
        for (uint32_t e = 0; e < E32; ++e) {
            const double beta = betas[e < s->E ? e : 0];
            for (int c = 0; c < ncls; ++c)
                t64[(size_t)e * 3 + c] = threshold64(beta, 4.0 * (c + 1) * h.jabs, s->planes);
        }
        if (!s->d_t64) {
            CUDA_TRY(ctx, dev_alloc(&s->d_t64, t64.size()));
            CUDA_TRY(ctx, dev_alloc(&s->d_tplane, (size_t)W * 3 * 8));
            CUDA_TRY(ctx, dev_alloc(&s->d_tlow, (size_t)E32 * 3));
        }
        CUDA_TRY(ctx, cudaMemcpyAsync(s->d_t64, t64.data(), t64.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        // no host wait: cudaMemcpyAsync from pageable memory returns once the buffer is staged
        count_launch(s, launch_build_tables_stencil(s->d_t64, W, s->planes, s->d_tplane, s->d_tlow, ctx->stream));
        s->perbeta = true;
        return ISING_OK;
    }
    const size_t per = (size_t)(GEN_MAX_DEG + 1) * GEN_MAX_CLS;
    std::vector<unsigned long long> t64((size_t)E32 * per, 0ull);
    std::vector<uint32_t> slot(E32);
    if (s->beta_rows_planes != s->planes || s->beta_rows.size() > 65536) {
        s->beta_rows.clear();
        s->beta_rows_planes = s->planes;
    }
    for (uint32_t e = 0; e < E32; ++e) {
        slot[e] = e;
        const double beta = betas[e < s->E ? e : 0];  // padding bits: any valid beta
        uint64_t key;
        memcpy(&key, &beta, sizeof key);
        auto it = s->beta_rows.find(key);
        if (it == s->beta_rows.end()) {
            std::vector<unsigned long long> row(per, 0ull);
            for (uint32_t deg = 1; deg <= (uint32_t)GEN_MAX_DEG; ++deg) {
                const uint32_t cmin = deg / 2 + 1, ncls = deg - deg / 2;
                for (uint32_t j = 0; j < ncls; ++j) {
                    const int cls = 2 * (int)(cmin + j) - (int)deg;
                    row[(size_t)deg * GEN_MAX_CLS + j] = threshold64(beta, 2.0 * h.jabs * (double)cls, s->planes);
                }
            }
            it = s->beta_rows.emplace(key, std::move(row)).first;
        }
        memcpy(&t64[(size_t)e * per], it->second.data(), per * sizeof(unsigned long long));
    }
    if (!s->d_t64) {
        CUDA_TRY(ctx, dev_alloc(&s->d_t64, t64.size()));
        CUDA_TRY(ctx, dev_alloc(&s->d_slot, (size_t)E32));
        CUDA_TRY(ctx, dev_alloc(&s->d_tplane, (size_t)(GEN_MAX_DEG + 1) * W * GEN_MAX_CLS * 8));
        CUDA_TRY(ctx, dev_alloc(&s->d_tlow, (size_t)(GEN_MAX_DEG + 1) * E32 * GEN_MAX_CLS));
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(s->d_t64, t64.data(), t64.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(s->d_slot, slot.data(), slot.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    count_launch(s, launch_build_tables(s->d_t64, s->d_slot, W, s->planes, s->d_tplane, s->d_tlow,
                                        ctx->stream));
    // no host wait: cudaMemcpyAsync from pageable memory returns once t64 / slot are staged
    s->perbeta = true;
    return ISING_OK;
}

struct ising_pt {
    ising_ctx* ctx = nullptr;
    const ising_graph* g = nullptr;
    ising_sim* sim = nullptr;
    uint64_t R = 0, lo = 0, hi = 0;   // this rank owns configurations [lo, hi)
    uint64_t word_lo = 0;             // first replica word held locally
    std::vector<double> betas;        // by slot
    std::vector<uint32_t> slot_of_cfg, cfg_of_slot;
    uint64_t seed = 0, swap_step = 0, total_swaps = 0;
};

// betas of the locally held replica bits from the slot permutation
static int pt_push_betas(ising_pt* pt) {
    const uint64_t E = pt->sim->E;
    std::vector<double> b(E);
    for (uint64_t e = 0; e < E; ++e) {
        const uint64_t cfg = pt->word_lo * 32 + e;
        b[e] = pt->betas[cfg < pt->R ? pt->slot_of_cfg[cfg] : 0];
    }
    return ising_sim_set_betas(pt->sim, b.data());
}

extern "C" int ising_pt_create(ising_ctx* ctx, const ising_graph* g, const double* betas,
                               uint64_t nbetas, uint64_t cfg_lo, uint64_t cfg_hi, uint64_t seed,
                               ising_pt** out) {
    CtxLock _lk(ctx);
    if (!ctx || !g || !betas || !out) return fail(ctx, ISING_E_INVALID, "ctx/graph/betas/out is NULL");
    *out = nullptr;
    if (nbetas == 0 || cfg_lo >= cfg_hi || cfg_hi > nbetas)
        return fail(ctx, ISING_E_INVALID, "need 0 <= cfg_lo < cfg_hi <= nbetas");
    std::unique_ptr<ising_pt> pt(new ising_pt);
    pt->ctx = ctx;
    pt->g = g;
    pt->R = nbetas;
    pt->lo = cfg_lo;
    pt->hi = cfg_hi;
    pt->seed = seed;
    pt->betas.assign(betas, betas + nbetas);
    pt->slot_of_cfg.resize(nbetas);
    pt->cfg_of_slot.resize(nbetas);
    for (uint64_t r = 0; r < nbetas; ++r) pt->slot_of_cfg[r] = pt->cfg_of_slot[r] = (uint32_t)r;
    // whole replica words: configuration c always lives at bit c%32 of global word c/32, so a
    // sharded run draws exactly the random numbers of the unsharded one
    pt->word_lo = cfg_lo / 32;
    const uint64_t word_hi = (cfg_hi + 31) / 32;
    const uint64_t E = std::min<uint64_t>((word_hi - pt->word_lo) * 32, nbetas - pt->word_lo * 32);
    // lattices temper on the checkerboard kernels, other graphs on the colour x degree kernels
    int rc = ising_sim_create_ex(ctx, g, E, seed, pt->word_lo * 32, 0u, &pt->sim);
    if (rc) return rc;
    rc = pt_push_betas(pt.get());
    if (rc) { ising_sim_destroy(pt->sim); return rc; }
    *out = pt.release();
    return ISING_OK;
}

extern "C" void ising_pt_destroy(ising_pt* pt) {
    CtxLock _lk(pt ? pt->ctx : nullptr);
    if (!pt) return;
    ising_sim_destroy(pt->sim);
    delete pt;
}

extern "C" int ising_pt_configure(ising_pt* pt, int planes, int rounds) {
    CtxLock _lk(pt ? pt->ctx : nullptr);
    if (!pt) return fail(nullptr, ISING_E_INVALID, "pt is NULL");
    const int rc = ising_sim_configure(pt->sim, planes, rounds);
    return rc ? rc : pt_push_betas(pt);
}

extern "C" int ising_pt_sweeps(ising_pt* pt, uint64_t t, double* local_energies) {
    CtxLock _lk(pt ? pt->ctx : nullptr);
    if (!pt) return fail(nullptr, ISING_E_INVALID, "pt is NULL");
    if (!local_energies) return ising_sim_sweeps(pt->sim, nullptr, t, nullptr);
    // enqueue only: the energy read-back below is the one host wait of the swap cycle
    CUDA_TRY(pt->ctx, cudaSetDevice(pt->ctx->device));
    int rc = sim_enqueue_sweeps(pt->sim, nullptr, t);
    if (rc) return rc;
    std::vector<double> en(pt->sim->E);
    rc = ising_sim_get_energies(pt->sim, en.data());
    if (rc) return rc;
    for (uint64_t c = pt->lo; c < pt->hi; ++c) local_energies[c - pt->lo] = en[c - pt->word_lo * 32];
    return ISING_OK;
}

// One tempering step (the shape of TemperingContainer::parallel_tempering_step as driven from
// tempering.rs:191-194): even slot pairs (0,1),(2,3).. then odd pairs (1,2),(3,4)..; the pair
// (a, a+1) exchanges configurations with probability min(1, exp((b_a - b_{a+1})(E_a - E_{a+1}))).
// The uniform is Philox(seed; slot a, swap step), so every rank takes the same decisions from
// the all-gathered energies.  all_energies is indexed by CONFIGURATION.
extern "C" int ising_pt_decide_swaps(const double* betas, uint64_t R, const double* all_energies,
                                     uint64_t seed, uint64_t swap_step, uint32_t* slot_of_cfg,
                                     uint32_t* cfg_of_slot, uint64_t* nswaps) {
    if (!betas || !all_energies || !slot_of_cfg || !cfg_of_slot)
        return fail(nullptr, ISING_E_INVALID, "ising_pt_decide_swaps: NULL argument");
    uint64_t swaps = 0;
    for (int parity = 0; parity < 2; ++parity)
        for (uint64_t a = parity; a + 1 < R; a += 2) {
            const uint32_t ca = cfg_of_slot[a], cb = cfg_of_slot[a + 1];
            const double d = (betas[a] - betas[a + 1]) * (all_energies[ca] - all_energies[cb]);
            bool acc = true;
            if (d < 0.0) {
                const u32x4 r = philox4x32<10>((uint32_t)a, (uint32_t)parity, (uint32_t)swap_step,
                                               TAG_SWAP << 24, (uint32_t)seed, (uint32_t)(seed >> 32));
                const double uu = ((double)r.x + 0.5) * (1.0 / 4294967296.0);
                acc = uu < exp(d);
            }
            if (acc) {
                cfg_of_slot[a] = cb;
                cfg_of_slot[a + 1] = ca;
                slot_of_cfg[cb] = (uint32_t)a;
                slot_of_cfg[ca] = (uint32_t)(a + 1);
                ++swaps;
            }
        }
    if (nswaps) *nswaps = swaps;
    return ISING_OK;
}

extern "C" int ising_pt_swap_step(ising_pt* pt, const double* all_energies) {
    CtxLock _lk(pt ? pt->ctx : nullptr);
    if (!pt || !all_energies) return fail(pt ? pt->ctx : nullptr, ISING_E_INVALID, "pt/energies is NULL");
    uint64_t swaps = 0;
    const int rc = ising_pt_decide_swaps(pt->betas.data(), pt->R, all_energies, pt->seed, pt->swap_step,
                                         pt->slot_of_cfg.data(), pt->cfg_of_slot.data(), &swaps);
    if (rc) return rc;
    pt->total_swaps += swaps;
    pt->swap_step++;
    return pt_push_betas(pt);
}

extern "C" int ising_pt_get_slots(const ising_pt* pt, uint32_t* slot_of_config) {
    CtxLock _lk(pt ? pt->ctx : nullptr);
    if (!pt || !slot_of_config) return fail(nullptr, ISING_E_INVALID, "pt/out is NULL");
    for (uint64_t c = 0; c < pt->R; ++c) slot_of_config[c] = pt->slot_of_cfg[c];
    return ISING_OK;
}

extern "C" int ising_pt_get_local_states(ising_pt* pt, uint8_t* states) {
    CtxLock _lk(pt ? pt->ctx : nullptr);
    if (!pt || !states) return fail(pt ? pt->ctx : nullptr, ISING_E_INVALID, "pt/states is NULL");
    const uint64_t N = pt->g->h.nvars;
    std::vector<uint8_t> all((size_t)pt->sim->E * N);
    const int rc = ising_sim_get_states(pt->sim, all.data());
    if (rc) return rc;
    for (uint64_t c = pt->lo; c < pt->hi; ++c)
        memcpy(states + (c - pt->lo) * N, all.data() + (c - pt->word_lo * 32) * N, N);
    return ISING_OK;
}

// checkpoint support: the sim behind the ladder, and the permutation / counters
extern "C" int ising_pt_get_sim(ising_pt* pt, ising_sim** out) {
    CtxLock _lk(pt ? pt->ctx : nullptr);
    if (!pt || !out) return fail(nullptr, ISING_E_INVALID, "pt/out is NULL");
    *out = pt->sim;
    return ISING_OK;
}

extern "C" int ising_pt_get_counters(const ising_pt* pt, uint64_t* swap_step, uint64_t* total_swaps) {
    CtxLock _lk(pt ? pt->ctx : nullptr);
    if (!pt || !swap_step || !total_swaps) return fail(nullptr, ISING_E_INVALID, "pt/out is NULL");
    *swap_step = pt->swap_step;
    *total_swaps = pt->total_swaps;
    return ISING_OK;
}

extern "C" int ising_pt_restore(ising_pt* pt, const uint32_t* slot_of_config, uint64_t swap_step,
                                uint64_t total_swaps) {
    CtxLock _lk(pt ? pt->ctx : nullptr);
    if (!pt || !slot_of_config) return fail(pt ? pt->ctx : nullptr, ISING_E_INVALID, "pt/slots is NULL");
    std::vector<uint8_t> seen(pt->R, 0);
    for (uint64_t c = 0; c < pt->R; ++c) {
        if (slot_of_config[c] >= pt->R || seen[slot_of_config[c]])
            return fail(pt->ctx, ISING_E_INVALID, "slot_of_config is not a permutation");
        seen[slot_of_config[c]] = 1;
    }
    for (uint64_t c = 0; c < pt->R; ++c) {
        pt->slot_of_cfg[c] = slot_of_config[c];
        pt->cfg_of_slot[slot_of_config[c]] = (uint32_t)c;
    }
    pt->swap_step = swap_step;
    pt->total_swaps = total_swaps;
    return pt_push_betas(pt);
}

extern "C" int ising_pt_total_swaps(const ising_pt* pt, uint64_t* out) {
    CtxLock _lk(pt ? pt->ctx : nullptr);
    if (!pt || !out) return fail(nullptr, ISING_E_INVALID, "pt/out is NULL");
    *out = pt->total_swaps;
    return ISING_OK;
}

// LatticeTempering::qmc_timesteps_sample, tempering.rs:156-222, single rank (cfg range = all):
// run min(to_sample, to_swap, remaining) -> swap step -> sample; states[R, n_s, nvars] holds
// "the configuration currently at beta_r", energies[R] = sum(E_r after chunk * chunk) / timesteps.
extern "C" int ising_pt_timesteps_sample(ising_pt* pt, uint64_t timesteps, uint64_t replica_swap_freq,
                                         uint64_t sampling_freq, uint8_t* states, double* energies) {
    CtxLock _lk(pt ? pt->ctx : nullptr);
    if (!pt || !energies) return fail(pt ? pt->ctx : nullptr, ISING_E_INVALID, "pt/energies is NULL");
    if (pt->lo != 0 || pt->hi != pt->R)
        return fail(pt->ctx, ISING_E_INVALID, "ising_pt_timesteps_sample needs all configurations on this rank");
    if (replica_swap_freq == 0 || sampling_freq == 0)
        return fail(pt->ctx, ISING_E_INVALID,
                    "replica_swap_freq and sampling_freq must be > 0 (the reference loops forever on 0)");
    const uint64_t R = pt->R, N = pt->g->h.nvars, ns = timesteps / sampling_freq;
    if (ns && !states) return fail(pt->ctx, ISING_E_INVALID, "states is NULL");
    std::vector<double> acc(R, 0.0), en(R);
    std::vector<uint8_t> local;
    uint64_t remaining = timesteps, to_swap = replica_swap_freq, to_sample = sampling_freq, k = 0;
    while (remaining > 0) {
        const uint64_t t = std::min(std::min(to_sample, to_swap), remaining);
        int rc = ising_pt_sweeps(pt, t, en.data());
        if (rc) return rc;
        for (uint64_t slot = 0; slot < R; ++slot) acc[slot] += en[pt->cfg_of_slot[slot]] * (double)t;
        to_sample -= t; to_swap -= t; remaining -= t;
        if (to_swap == 0) {
            rc = ising_pt_swap_step(pt, en.data());
            if (rc) return rc;
            to_swap = replica_swap_freq;
        }
        if (to_sample == 0) {
            if (k < ns) {
                local.resize((size_t)R * N);
                rc = ising_pt_get_local_states(pt, local.data());
                if (rc) return rc;
                for (uint64_t slot = 0; slot < R; ++slot)
                    memcpy(states + (slot * ns + k) * N, local.data() + (size_t)pt->cfg_of_slot[slot] * N, N);
            }
            ++k;
            to_sample = sampling_freq;
        }
    }
    for (uint64_t slot = 0; slot < R; ++slot) energies[slot] = acc[slot] / (double)timesteps;
    return ISING_OK;
}
