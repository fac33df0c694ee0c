// Initial particle loading.  Replaces MaxwellianParticleInitializer::loadParticles
// (src/core/data/ions/particle_initializers/maxwellian_particle_initializer.hpp:138-199, .cpp:12-84).
//
// Two entry points:
//  * phb_maxwellian_load_host — "parity mode": the reference's sequential algorithm on the host with the same
//    standard-library generators (std::mt19937_64 seeded with the population seed, a fresh
//    std::normal_distribution per velocity component and particle, std::uniform_real_distribution(0, 1-eps) for
//    the deltas, velocities drawn before deltas), cells walked in the row-major order of
//    layout.indices(AMRBox) (gridlayout.hpp:219-226).  Built against the same libstdc++ the reference would
//    be, it yields the reference's particles bit for bit (tests/test_frontend.py checks it against the
//    reference initializer compiled in place).  Initialisation only — nothing on the step path calls it.
//  * phb_maxwellian_load — the B200 loader: one thread per particle, counter-based Philox4x32-10 keyed by the
//    seed and indexed by (global cell, particle-in-cell) (so the result does not depend on the patch
//    decomposition or the launch shape), Box-Muller normals, written straight into the device-resident SoA
//    store in cell order (the store is born binned: d_first doubles as phb_bin's cell_start for the domain
//    keys).  It cannot reproduce the mt19937_64 stream; it is validated statistically
//    (tests/simulator/initialize/density_check.py style: density, bulk velocity and thermal spread per cell).
#include "common.cuh"

#include <cmath>
#include <limits>
#include <optional>
#include <random>
#include <vector>

namespace phb
{
// ---------------------------------------------------------------------------------------- host, parity mode
static void local_magnetic_basis(double const B[3], double basis[3][3])
{
    // localMagneticBasis, maxwellian_particle_initializer.cpp:44-84
    double const b2 = std::sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
    if (b2 < 1e-8)
    {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                basis[i][j] = i == j ? 1.0 : 0.0;
        return;
    }
    for (int j = 0; j < 3; ++j)
        basis[0][j] = B[j] / b2;
    basis[1][0] = B[2] - B[1];
    basis[1][1] = B[0] - B[2];
    basis[1][2] = B[1] - B[0];
    double const n1 = std::sqrt(basis[1][0] * basis[1][0] + basis[1][1] * basis[1][1] + basis[1][2] * basis[1][2]);
    for (int j = 0; j < 3; ++j)
        basis[1][j] /= n1;
    basis[2][0] = basis[0][1] * basis[1][2] - basis[0][2] * basis[1][1];
    basis[2][1] = basis[0][2] * basis[1][0] - basis[0][0] * basis[1][2];
    basis[2][2] = basis[0][0] * basis[1][1] - basis[0][1] * basis[1][0];
}

// ---------------------------------------------------------------------------------------- device loader
struct Philox
{
    // Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0,k1)
    static __host__ __device__ __forceinline__ void round(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
    {
        constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        uint64_t const p0 = uint64_t(M0) * c[0], p1 = uint64_t(M1) * c[2];
        uint32_t const n0 = uint32_t(p1 >> 32) ^ c[1] ^ k0, n1 = uint32_t(p1);
        uint32_t const n2 = uint32_t(p0 >> 32) ^ c[3] ^ k1, n3 = uint32_t(p0);
        c[0] = n0, c[1] = n1, c[2] = n2, c[3] = n3;
    }
    static __host__ __device__ __forceinline__ void run(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
    {
        constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
        for (int r = 0; r < 10; ++r)
        {
            round(c, k0, k1);
            k0 += W0;
            k1 += W1;
        }
    }
};

// two uniform doubles in [0,1) with 53 random bits each from one Philox block
__device__ __forceinline__ void uniform2(uint64_t index, uint32_t block, uint64_t seed, double& u0, double& u1)
{
    uint32_t c[4] = {uint32_t(index), uint32_t(index >> 32), block, 0x50484152u /* "PHAR" */};
    Philox::run(c, uint32_t(seed), uint32_t(seed >> 32));
    uint64_t const a = (uint64_t(c[0]) << 32) | c[1], b = (uint64_t(c[2]) << 32) | c[3];
    u0 = double(a >> 11) * 0x1.0p-53;
    u1 = double(b >> 11) * 0x1.0p-53;
}

template<int DIM>
struct LoadParams
{
    DevLayout L;
    const double* n;
    const double* V[3];
    const double* Vth[3];
    const uint32_t* first; // first particle of each cell (ncell + 1 entries)
    PartView out;
    size_t out_first;       // offset into the store
    uint64_t seed;
    int domain_cells[3];    // level domain extent: the counter is keyed by the GLOBAL cell, not the patch-local one
    double charge;
    uint32_t ppc;
    size_t ncell;
};

template<int DIM>
__global__ void __launch_bounds__(256) maxwellian_load_kernel(const __grid_constant__ LoadParams<DIM> A)
{
    size_t const t = size_t(blockIdx.x) * blockDim.x + threadIdx.x; // slot = cell * ppc + ipart
    if (t >= A.ncell * A.ppc)
        return;
    size_t const cell    = t / A.ppc;
    uint32_t const ipart = uint32_t(t % A.ppc);
    uint32_t const f0 = A.first[cell], f1 = A.first[cell + 1];
    if (ipart >= f1 - f0) // cell below the density cut-off
        return;
    size_t const p = A.out_first + f0 + ipart;
    // cell -> AMR index (row-major over the patch box)
    size_t rem = cell;
    int ic[3] = {0, 0, 0};
#pragma unroll
    for (int d = DIM - 1; d >= 0; --d)
    {
        ic[d] = int(rem % size_t(A.L.ncells[d])) + A.L.amr_lower[d];
        rem /= size_t(A.L.ncells[d]);
    }
    uint64_t gcell = 0; // row-major index of the cell in the level's (periodic) domain box
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        int const ext = A.domain_cells[d];
        gcell         = gcell * uint64_t(ext) + uint64_t(((ic[d] % ext) + ext) % ext);
    }
    uint64_t const gidx = gcell * A.ppc + ipart;
    double u[8];
    uniform2(gidx, 0, A.seed, u[0], u[1]);
    uniform2(gidx, 1, A.seed, u[2], u[3]);
    uniform2(gidx, 2, A.seed, u[4], u[5]);
    uniform2(gidx, 3, A.seed, u[6], u[7]);
    // Box-Muller: (u0,u1) -> z0,z1 ; (u2,u3) -> z2
    double s0, c0, s1, c1;
    sincospi(2. * u[1], &s0, &c0);
    sincospi(2. * u[3], &s1, &c1);
    double const r0 = sqrt(-2. * log(1. - u[0])), r1 = sqrt(-2. * log(1. - u[2]));
    double const z[3] = {r0 * c0, r0 * s0, r1 * c1};
#pragma unroll
    for (int c = 0; c < 3; ++c)
        A.out.v[c][p] = A.V[c][cell] + A.Vth[c][cell] * z[c];
#pragma unroll
    for (int d = 0; d < DIM; ++d)
    {
        A.out.icell[d][p] = ic[d];
        // same support as ParticleDeltaDistribution: [0, 1 - eps)  (particle.hpp:19-27)
        A.out.delta[d][p] = u[4 + d] * (1. - 2.220446049250313e-16);
    }
    A.out.weight[p] = A.n[cell] / double(A.ppc);
    A.out.charge[p] = A.charge;
}

template<int DIM>
int load_dim(phb_ctx* ctx, const phb_layout* L, const double* d_n, const phb_vecfield* d_V, const phb_vecfield* d_Vth,
             const uint32_t* d_first, double charge, uint32_t ppc, uint64_t seed, const uint32_t* domain_cells,
             phb_particles* out, size_t out_first)
{
    LoadParams<DIM> A;
    A.L = make_dev_layout(*L);
    A.n = d_n;
    for (int c = 0; c < 3; ++c)
    {
        A.V[c]   = d_V->comp[c];
        A.Vth[c] = d_Vth->comp[c];
    }
    A.first     = d_first;
    A.out       = make_part(*out);
    A.out_first = out_first;
    A.seed      = seed;
    for (int d = 0; d < 3; ++d)
        A.domain_cells[d] = d < DIM ? int(domain_cells ? domain_cells[d] : L->ncells[d]) : 1;
    A.charge    = charge;
    A.ppc       = ppc;
    A.ncell     = 1;
    for (int d = 0; d < DIM; ++d)
        A.ncell *= size_t(L->ncells[d]);
    size_t const slots = A.ncell * ppc;
    if (slots == 0)
        return PHB_OK;
    maxwellian_load_kernel<DIM><<<unsigned((slots + 255) / 256), 256, 0, ctx->stream>>>(A);
    PHB_LAUNCH_CHECK(ctx);
    return PHB_OK;
}
} // namespace phb

extern "C" int phb_maxwellian_load_host(const phb_layout* L, const double* h_n, const double* const* h_V,
                                        const double* const* h_Vth, const double* const* h_B, double charge,
                                        uint32_t ppc, int has_seed, size_t seed, double density_cut_off, int* h_icell,
                                        double* h_delta, double* h_weight, double* h_charge, double* h_v, size_t capacity,
                                        size_t* h_count)
{
    if (!L || L->dim < 1 || L->dim > 3 || !h_n || !h_V || !h_Vth || !h_count)
        return PHB_ERR_INVALID;
    int const dim = L->dim;
    size_t ncell  = 1;
    for (int d = 0; d < dim; ++d)
        ncell *= size_t(L->ncells[d]);
    // getRNG, maxwellian_particle_initializer.hpp:71-81
    std::mt19937_64 gen;
    if (has_seed)
        gen = std::mt19937_64(seed);
    else
    {
        std::random_device rd;
        std::seed_seq sq{rd(), rd(), rd(), rd(), rd(), rd(), rd(), rd()};
        gen = std::mt19937_64(sq);
    }
    std::uniform_real_distribution<double> delta_dist{0, 1. - std::numeric_limits<double>::epsilon()};
    size_t count = 0;
    for (size_t cell = 0; cell < ncell; ++cell)
    {
        if (h_n[cell] < density_cut_off)
            continue;
        if (count + ppc > capacity)
            return PHB_ERR_CAPACITY;
        double const cell_weight = h_n[cell] / ppc;
        int ic[3] = {0, 0, 0};
        size_t rem = cell;
        for (int d = dim - 1; d >= 0; --d)
        {
            ic[d] = int(rem % size_t(L->ncells[d])) + L->amr_lower[d];
            rem /= size_t(L->ncells[d]);
        }
        double basis[3][3];
        if (h_B)
        {
            double const B[3] = {h_B[0][cell], h_B[1][cell], h_B[2][cell]};
            phb::local_magnetic_basis(B, basis);
        }
        for (uint32_t ip = 0; ip < ppc; ++ip, ++count)
        {
            double v[3];
            for (int c = 0; c < 3; ++c) // maxwellianVelocity: a fresh distribution per component and particle
            {
                std::normal_distribution<> maxwell(h_V[c][cell], h_Vth[c][cell]);
                v[c] = maxwell(gen);
            }
            if (h_B) // basisTransform, .cpp:28-40
            {
                double w[3];
                for (int c = 0; c < 3; ++c)
                    w[c] = basis[0][c] * v[0] + basis[1][c] * v[1] + basis[2][c] * v[2];
                for (int c = 0; c < 3; ++c)
                    v[c] = w[c];
            }
            for (int c = 0; c < 3; ++c)
                h_v[count * 3 + c] = v[c];
            for (int d = 0; d < dim; ++d)
            {
                h_icell[count * dim + d] = ic[d];
                h_delta[count * dim + d] = delta_dist(gen);
            }
            h_weight[count] = cell_weight;
            h_charge[count] = charge;
        }
    }
    *h_count = count;
    return PHB_OK;
}

extern "C" int phb_maxwellian_load(phb_ctx* ctx, const phb_layout* L, const double* d_n, const phb_vecfield* d_V,
                                   const phb_vecfield* d_Vth, const uint32_t* d_first, size_t h_total, double charge,
                                   uint32_t ppc, uint64_t seed, const uint32_t* h_domain_cells, phb_particles* out)
{
    if (!phb::valid_layout(ctx, L) || !d_n || !d_V || !d_Vth || !d_first || !out || ppc == 0)
        return phb::set_error(ctx, PHB_ERR_INVALID, "phb_maxwellian_load: invalid argument");
    if (out->n + h_total > out->capacity)
        return phb::set_error(ctx, PHB_ERR_CAPACITY, "phb_maxwellian_load: store capacity exceeded");
    int rc;
    switch (L->dim)
    {
        case 1: rc = phb::load_dim<1>(ctx, L, d_n, d_V, d_Vth, d_first, charge, ppc, seed, h_domain_cells, out, out->n); break;
        case 2: rc = phb::load_dim<2>(ctx, L, d_n, d_V, d_Vth, d_first, charge, ppc, seed, h_domain_cells, out, out->n); break;
        default: rc = phb::load_dim<3>(ctx, L, d_n, d_V, d_Vth, d_first, charge, ppc, seed, h_domain_cells, out, out->n); break;
    }
    if (rc == PHB_OK)
        out->n += h_total;
    return rc;
}
