// Host-side graph compiler (see graph.h).
#include "graph.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <queue>

#include "../../include/ising_b200.h"
#include "philox.h"

namespace ising {

// ------------------------------------------------------------------------------------------
// seeds: SmallRng (xoshiro256++ seeded through SplitMix64), src/lattice.rs:83-91
// ------------------------------------------------------------------------------------------
namespace {
struct Xoshiro256pp {
    uint64_t s[4];
    explicit Xoshiro256pp(uint64_t seed) {
        for (auto& w : s) {
            seed += 0x9e3779b97f4a7c15ULL;
            uint64_t z = seed;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            w = z ^ (z >> 31);
        }
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t out = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return out;
    }
};
}  // namespace

void make_seeds(uint64_t seed_gen, uint64_t n, uint64_t* out) {
    Xoshiro256pp rng(seed_gen);
    for (uint64_t i = 0; i < n; ++i) out[i] = rng.next();
}

// ------------------------------------------------------------------------------------------
// annealing schedule, src/lattice.rs:320-334 and :357-365 (same code again at :406-420, :445-451)
// ------------------------------------------------------------------------------------------
bool schedule_betas(const uint64_t* st, const double* sb, uint64_t n, uint64_t timesteps,
                    bool linear, double* out) {
    std::vector<std::pair<uint64_t, double>> stops;
    stops.reserve(n + 3);
    for (uint64_t k = 0; k < n; ++k) stops.emplace_back(st[k], sb[k]);
    std::stable_sort(stops.begin(), stops.end(),
                     [](const std::pair<uint64_t, double>& x, const std::pair<uint64_t, double>& y) {
                         return x.first < y.first;
                     });
    if (stops.empty()) {
        stops.emplace_back(0, 1.0);
        stops.emplace_back(timesteps, 1.0);
    }
    if (stops.front().first > 0) stops.insert(stops.begin(), {0, stops.front().second});
    // The reference's closure captures this `i` and never uses its own timestep argument.
    const uint64_t i_captured = stops.back().first;
    if (i_captured < timesteps) stops.emplace_back(timesteps, stops.back().second);
    if (stops.size() < 2) return false;
    size_t idx = 0;
    for (uint64_t t = 0; t < timesteps; ++t) {
        const uint64_t i = linear ? t : i_captured;
        while (idx + 2 < stops.size() && i > stops[idx + 1].first) ++idx;
        const uint64_t ia = stops[idx].first, ib = stops[idx + 1].first;
        const double va = stops[idx].second, vb = stops[idx + 1].second;
        out[t] = (vb - va) * ((double)(i - ia) / (double)(ib - ia)) + va;
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// HostGraph
// ------------------------------------------------------------------------------------------
static inline uint64_t torus_fwd(const uint64_t dims[3], uint64_t n, int d) {
    const uint64_t Lx = dims[0], Ly = dims[1], Lz = dims[2];
    const uint64_t x = n % Lx, y = (n / Lx) % Ly, z = n / (Lx * Ly);
    if (d == 0) return (x + 1 == Lx ? 0 : x + 1) + Lx * (y + Ly * z);
    if (d == 1) return x + Lx * ((y + 1 == Ly ? 0 : y + 1) + Ly * z);
    return x + Lx * (y + Ly * (z + 1 == Lz ? 0 : z + 1));
}

void HostGraph::edge_at(uint64_t e, uint64_t* a, uint64_t* b, double* j) const {
    if (!implicit_edges) {
        *a = ea[e];
        *b = eb[e];
        *j = ej[e];
        return;
    }
    const int dim = kind == ISING_KIND_STENCIL3D ? 3 : 2;
    const uint64_t n = e / dim;
    const int d = (int)(e % dim);
    *a = n;
    *b = torus_fwd(dims, n, d);
    *j = ((fwd_sign[n] >> d) & 1) ? jabs : -jabs;
}

void HostGraph::build_csr() {
    if (csr_built) return;
    row.assign(nvars + 1, 0);
    for (uint64_t e = 0; e < nedges; ++e) {
        uint64_t a, b;
        double j;
        edge_at(e, &a, &b, &j);
        row[a + 1]++;
        row[b + 1]++;
    }
    for (uint64_t i = 0; i < nvars; ++i) row[i + 1] += row[i];
    nbr.assign(2 * nedges, 0);
    jv.assign(2 * nedges, 0.0);
    std::vector<uint64_t> fill(row.begin(), row.end() - 1);
    for (uint64_t e = 0; e < nedges; ++e) {
        uint64_t a, b;
        double j;
        edge_at(e, &a, &b, &j);
        nbr[fill[a]] = (uint32_t)b;
        jv[fill[a]++] = j;
        nbr[fill[b]] = (uint32_t)a;
        jv[fill[b]++] = j;
    }
    // stable sort of every adjacency list by neighbour index (insertion order = edge order)
    max_degree = 0;
    std::vector<uint32_t> perm;
    std::vector<uint32_t> tn;
    std::vector<double> tj;
    for (uint64_t i = 0; i < nvars; ++i) {
        const uint64_t lo = row[i], hi = row[i + 1];
        const uint64_t deg = hi - lo;
        max_degree = std::max<int>(max_degree, (int)deg);
        bool sorted = true;
        for (uint64_t k = lo + 1; k < hi; ++k)
            if (nbr[k - 1] > nbr[k]) { sorted = false; break; }
        if (sorted) continue;
        perm.resize(deg);
        std::iota(perm.begin(), perm.end(), 0u);
        std::stable_sort(perm.begin(), perm.end(),
                         [&](uint32_t x, uint32_t y) { return nbr[lo + x] < nbr[lo + y]; });
        tn.resize(deg);
        tj.resize(deg);
        for (uint64_t k = 0; k < deg; ++k) { tn[k] = nbr[lo + perm[k]]; tj[k] = jv[lo + perm[k]]; }
        for (uint64_t k = 0; k < deg; ++k) { nbr[lo + k] = tn[k]; jv[lo + k] = tj[k]; }
    }
    csr_built = true;
}

uint32_t HostGraph::color_of(uint64_t n) const {
    if (kind == ISING_KIND_GENERAL) return color[n];
    const uint64_t x = n % dims[0], y = (n / dims[0]) % dims[1], z = n / (dims[0] * dims[1]);
    return (uint32_t)((x + y + z) & 1);
}

// Try to read the edge list as a row-major periodic square / cubic lattice with even extents
// >= 4 and every bond present exactly once.  Fills kind/dims/fwd_sign on success.
static bool recognise_torus(HostGraph* g) {
    const uint64_t N = g->nvars;
    if (!g->integer_classes || N < 16 || N > 0xFFFFFFFFull) return false;
    const uint64_t deg0 = g->row[1] - g->row[0];
    if (deg0 != 4 && deg0 != 6) return false;
    const int dim = (int)(deg0 / 2);
    if (g->nedges != (uint64_t)dim * N) return false;
    const uint32_t* nb = &g->nbr[g->row[0]];
    uint64_t Lx, Ly, Lz = 1;
    if (dim == 2) {
        Lx = nb[2];
        if (Lx < 4 || N % Lx) return false;
        Ly = N / Lx;
        if (!(nb[0] == 1 && nb[1] == Lx - 1 && nb[3] == N - Lx)) return false;
    } else {
        Lx = nb[2];
        const uint64_t LxLy = nb[4];
        if (Lx < 4 || LxLy % Lx || N % LxLy) return false;
        Ly = LxLy / Lx;
        Lz = N / LxLy;
        if (!(nb[0] == 1 && nb[1] == Lx - 1 && nb[3] == Lx * (Ly - 1) && nb[5] == N - LxLy))
            return false;
    }
    if (Lx < 4 || Ly < 4 || (dim == 3 && Lz < 4)) return false;
    if ((Lx & 1) || (Ly & 1) || (dim == 3 && (Lz & 1))) return false;  // odd: not bipartite
    const uint64_t dims[3] = {Lx, Ly, Lz};
    std::vector<uint8_t> sign(N, 0), seen(N, 0);
    for (uint64_t e = 0; e < g->nedges; ++e) {
        uint64_t a = g->ea[e], b = g->eb[e];
        int d = -1;
        uint64_t n = 0;
        for (int dd = 0; dd < dim; ++dd) {
            if (torus_fwd(dims, a, dd) == b) { d = dd; n = a; break; }
            if (torus_fwd(dims, b, dd) == a) { d = dd; n = b; break; }
        }
        if (d < 0 || ((seen[n] >> d) & 1)) return false;
        seen[n] |= (uint8_t)(1u << d);
        if (g->ej[e] > 0) sign[n] |= (uint8_t)(1u << d);
    }
    const uint8_t full = (uint8_t)((1u << dim) - 1);
    for (uint64_t n = 0; n < N; ++n)
        if (seen[n] != full) return false;
    g->kind = dim == 2 ? ISING_KIND_STENCIL2D : ISING_KIND_STENCIL3D;
    g->dims[0] = Lx; g->dims[1] = Ly; g->dims[2] = Lz;
    g->fwd_sign.swap(sign);
    return true;
}

static void classify_signs(HostGraph* g) {
    const int dim = g->kind == ISING_KIND_STENCIL3D ? 3 : 2;
    const uint8_t full = (uint8_t)((1u << dim) - 1);
    bool all0 = true, all1 = true;
    for (uint8_t s : g->fwd_sign) {
        all0 &= (s == 0);
        all1 &= (s == full);
    }
    g->uniform_sign = all0 || all1;
    g->uniform_antiferro = all1 && !all0;
}

// Bipartite BFS 2-colouring if possible, else greedy in order of decreasing degree.
static void colour_graph(HostGraph* g) {
    const uint64_t N = g->nvars;
    g->color.assign(N, 0xFFFFFFFFu);
    bool bipartite = true;
    std::vector<uint64_t> stack;
    for (uint64_t s = 0; s < N && bipartite; ++s) {
        if (g->color[s] != 0xFFFFFFFFu) continue;
        g->color[s] = 0;
        stack.push_back(s);
        while (!stack.empty() && bipartite) {
            const uint64_t u = stack.back();
            stack.pop_back();
            for (uint64_t k = g->row[u]; k < g->row[u + 1]; ++k) {
                const uint32_t v = g->nbr[k];
                if (v == u) { bipartite = false; break; }
                if (g->color[v] == 0xFFFFFFFFu) {
                    g->color[v] = g->color[u] ^ 1u;
                    stack.push_back(v);
                } else if (g->color[v] == g->color[u]) {
                    bipartite = false;
                    break;
                }
            }
        }
    }
    if (bipartite) {
        g->ncolors = 1;
        for (uint32_t c : g->color) if (c == 1) { g->ncolors = 2; break; }
        return;
    }
    std::vector<uint64_t> order(N);
    std::iota(order.begin(), order.end(), 0ull);
    std::stable_sort(order.begin(), order.end(), [&](uint64_t x, uint64_t y) {
        return (g->row[x + 1] - g->row[x]) > (g->row[y + 1] - g->row[y]);
    });
    std::fill(g->color.begin(), g->color.end(), 0xFFFFFFFFu);
    std::vector<uint64_t> mark(g->max_degree + 2, (uint64_t)-1);
    std::vector<uint64_t> class_size;
    int ncol = 0;
    for (uint64_t u : order) {
        for (uint64_t k = g->row[u]; k < g->row[u + 1]; ++k) {
            const uint32_t c = g->color[g->nbr[k]];
            if (c != 0xFFFFFFFFu && c < mark.size()) mark[c] = u;
        }
        // smallest-population admissible colour keeps the classes balanced
        int best = -1;
        for (int c = 0; c < ncol; ++c)
            if (mark[c] != u && (best < 0 || class_size[c] < class_size[best])) best = c;
        if (best < 0) {
            best = ncol++;
            class_size.push_back(0);
        }
        g->color[u] = (uint32_t)best;
        class_size[best]++;
    }
    g->ncolors = ncol;
}

std::string compile_from_edges(uint64_t nvars, uint64_t nedges, const uint64_t* a,
                               const uint64_t* b, const double* j, const double* biases,
                               HostGraph* g) {
    if (nedges == 0) return "Must supply some edges for graph";
    if (nvars == 0 || nvars > 0xFFFFFFFFull) return "nvars out of range";
    g->nvars = nvars;
    g->nedges = nedges;
    g->ea.assign(a, a + nedges);
    g->eb.assign(b, b + nedges);
    g->ej.assign(j, j + nedges);
    for (uint64_t e = 0; e < nedges; ++e) {
        if (a[e] >= nvars || b[e] >= nvars) return "edge endpoint out of range";
        if (a[e] == b[e]) return "self-loops cannot be coloured (edge with a == b)";
        if (!std::isfinite(j[e])) return "non-finite coupling";
    }
    g->bias.assign(nvars, 0.0);
    g->has_bias = false;
    if (biases)
        for (uint64_t i = 0; i < nvars; ++i) {
            g->bias[i] = biases[i];
            if (biases[i] != 0.0) g->has_bias = true;
        }
    g->build_csr();
    g->jabs = fabs(j[0]);
    g->integer_classes = !g->has_bias && g->jabs > 0.0;
    for (uint64_t e = 0; e < nedges && g->integer_classes; ++e)
        if (fabs(j[e]) != g->jabs) g->integer_classes = false;
    g->kind = ISING_KIND_GENERAL;
    if (recognise_torus(g)) {
        g->ncolors = 2;
        classify_signs(g);
    } else {
        colour_graph(g);
    }
    return "";
}

void strong_edge_colouring(HostGraph* g, EdgeClasses* out) {
    g->build_csr();
    const uint64_t N = g->nvars, M = g->nedges;
    // incident edge ids per site
    std::vector<uint64_t> irow(N + 1, 0);
    std::vector<uint32_t> ea(M), eb(M);
    std::vector<float> wj(M);
    double jmax = 0.0;
    for (uint64_t e = 0; e < M; ++e) {
        uint64_t a, b;
        double j;
        g->edge_at(e, &a, &b, &j);
        ea[e] = (uint32_t)a;
        eb[e] = (uint32_t)b;
        wj[e] = (float)fabs(j);
        jmax = std::max(jmax, fabs(j));
        irow[a + 1]++;
        irow[b + 1]++;
    }
    for (uint64_t i = 0; i < N; ++i) irow[i + 1] += irow[i];
    std::vector<uint32_t> inc(2 * M);
    {
        std::vector<uint64_t> fill(irow.begin(), irow.end() - 1);
        for (uint64_t e = 0; e < M; ++e) {
            inc[fill[ea[e]]++] = (uint32_t)e;
            inc[fill[eb[e]]++] = (uint32_t)e;
        }
    }
    std::vector<uint32_t> cls(M, 0xFFFFFFFFu);
    std::vector<uint64_t> stamp;   // stamp[c] == e + 1: class c is taken around edge e
    uint32_t ncls = 0;
    auto forbid_site = [&](uint64_t v, uint64_t e) {
        for (uint64_t k = irow[v]; k < irow[v + 1]; ++k) {
            const uint32_t c = cls[inc[k]];
            if (c != 0xFFFFFFFFu) stamp[c] = e + 1;
        }
    };
    for (uint64_t e = 0; e < M; ++e) {
        const uint32_t end[2] = {ea[e], eb[e]};
        for (int s = 0; s < 2; ++s) {
            forbid_site(end[s], e);
            for (uint64_t k = g->row[end[s]]; k < g->row[end[s] + 1]; ++k) forbid_site(g->nbr[k], e);
        }
        uint32_t c = 0;
        while (c < ncls && stamp[c] == e + 1) ++c;
        if (c == ncls) {
            ++ncls;
            stamp.push_back(0);
        }
        cls[e] = c;
    }
    out->off.assign(ncls + 1, 0);
    for (uint64_t e = 0; e < M; ++e) out->off[cls[e] + 1]++;
    for (uint32_t c = 0; c < ncls; ++c) out->off[c + 1] += out->off[c];
    out->ea.resize(M);
    out->eb.resize(M);
    out->eid.resize(M);
    out->wrel.resize(M);
    std::vector<uint32_t> fill(out->off.begin(), out->off.end() - 1);
    for (uint64_t e = 0; e < M; ++e) {
        const uint32_t p = fill[cls[e]]++;
        out->ea[p] = ea[e];
        out->eb[p] = eb[e];
        out->eid[p] = (uint32_t)e;
        out->wrel[p] = jmax > 0.0 ? (float)(wj[e] / jmax) : 1.f;
    }
}

std::string make_torus(int dim, const uint64_t* L, double j0, int pmj, uint64_t j_seed,
                       HostGraph* g) {
    if (dim != 2 && dim != 3) return "dim must be 2 or 3";
    uint64_t N = 1;
    for (int d = 0; d < dim; ++d) {
        if (L[d] < 4 || (L[d] & 1)) return "torus extents must be even and >= 4";
        N *= L[d];
        if (N > 0xFFFFFFFFull) return "torus too large for 32-bit site indices";
    }
    if (!(fabs(j0) > 0.0) || !std::isfinite(j0)) return "j0 must be finite and non-zero";
    g->nvars = N;
    g->nedges = (uint64_t)dim * N;
    g->implicit_edges = true;
    g->kind = dim == 2 ? ISING_KIND_STENCIL2D : ISING_KIND_STENCIL3D;
    g->dims[0] = L[0];
    g->dims[1] = L[1];
    g->dims[2] = dim == 3 ? L[2] : 1;
    g->jabs = fabs(j0);
    g->integer_classes = true;
    g->has_bias = false;
    g->max_degree = 2 * dim;
    g->ncolors = 2;
    g->fwd_sign.assign(N, 0);
    const uint8_t full = (uint8_t)((1u << dim) - 1);
    if (pmj) {
        const uint32_t k0 = (uint32_t)j_seed, k1 = (uint32_t)(j_seed >> 32);
        for (uint64_t n = 0; n < N; ++n) {
            // one Philox call per site, bit 0 of word d decides bond (n, d)
            const u32x4 r = philox4x32<10>((uint32_t)n, 0u, 0u, TAG_BOND << 24, k0, k1);
            uint8_t s = (uint8_t)((r.x & 1u) | ((r.y & 1u) << 1));
            if (dim == 3) s |= (uint8_t)((r.z & 1u) << 2);
            g->fwd_sign[n] = s;
        }
    } else if (j0 > 0) {
        std::fill(g->fwd_sign.begin(), g->fwd_sign.end(), full);
    }
    classify_signs(g);
    return "";
}

}  // namespace ising
