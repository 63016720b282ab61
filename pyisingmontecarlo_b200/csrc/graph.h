// Host-side graph compiler: edge list -> CSR, coupling classification, torus recognition,
// greedy colouring.  Replaces the per-experiment adjacency build of GraphState::new
// (reference call site src/lattice.rs:199) with one compile per Lattice.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace ising {

struct HostGraph {
    uint64_t nvars = 0;
    uint64_t nedges = 0;
    // explicit edge list (empty for implicit tori built by make_torus)
    std::vector<uint64_t> ea, eb;
    std::vector<double> ej;
    bool implicit_edges = false;

    std::vector<double> bias;  // nvars (all zero when !has_bias)
    bool has_bias = false;

    // CSR, neighbours ascending, ties in edge-list order (built on demand for implicit tori)
    bool csr_built = false;
    std::vector<uint64_t> row;
    std::vector<uint32_t> nbr;
    std::vector<double> jv;
    int max_degree = 0;

    bool integer_classes = false;  // all |J| equal (> 0) and no bias
    double jabs = 0.0;

    int kind = 0;  // ISING_KIND_*
    uint64_t dims[3] = {1, 1, 1};
    // stencil: bit d of fwd_sign[n] = 1 iff the bond n -> n + e_d has J > 0 (antiferro)
    std::vector<uint8_t> fwd_sign;
    bool uniform_sign = false;  // every bond has the same sign
    bool uniform_antiferro = false;

    int ncolors = 0;
    std::vector<uint32_t> color;  // general graphs only (stencil colour = parity of x+y+z)

    void build_csr();
    uint32_t color_of(uint64_t n) const;
    void edge_at(uint64_t e, uint64_t* a, uint64_t* b, double* j) const;
};

// Returns "" on success, else an error message.
std::string compile_from_edges(uint64_t nvars, uint64_t nedges, const uint64_t* a,
                               const uint64_t* b, const double* j, const double* biases,
                               HostGraph* out);
std::string make_torus(int dim, const uint64_t* L, double j0, int pmj, uint64_t j_seed,
                       HostGraph* out);

// rand 0.8 SmallRng restated for Lattice::make_seeds (src/lattice.rs:83-91)
void make_seeds(uint64_t seed_gen, uint64_t n, uint64_t* out);

// src/lattice.rs:320-334 + 357-365; linear = documented interpolation instead of quirk Q1.
// Returns false when the reference would index out of bounds.
bool schedule_betas(const uint64_t* st, const double* sb, uint64_t n, uint64_t timesteps,
                    bool linear, double* out);

}  // namespace ising
