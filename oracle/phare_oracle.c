/* TEST INFRASTRUCTURE — CPU oracle, see phare_oracle.h.  Plain C, scalar, one thread.
 * Every function restates the reference arithmetic in the reference's operation order; citations
 * are relative to the PHARE source tree.  Build: gcc -O2 -ffp-contract=off (no FMA contraction:
 * the reference's default build has no -march flag, so its products and sums are separately
 * rounded; fusing them changes the bits).
 */
#include "phare_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- index space: gridlayout.hpp:849-864,1349-1494 ; centerings gridlayout_hybrid_yee.hpp:54-85 */
enum { PRIMAL = 0, DUAL = 1 };
static const int CENTERING[PHB_NQTY][3] = {
    {PRIMAL, DUAL, DUAL},     {DUAL, PRIMAL, DUAL},     {DUAL, DUAL, PRIMAL},     /* Bx By Bz */
    {DUAL, PRIMAL, PRIMAL},   {PRIMAL, DUAL, PRIMAL},   {PRIMAL, PRIMAL, DUAL},   /* Ex Ey Ez */
    {DUAL, PRIMAL, PRIMAL},   {PRIMAL, DUAL, PRIMAL},   {PRIMAL, PRIMAL, DUAL},   /* Jx Jy Jz */
    {PRIMAL, PRIMAL, PRIMAL}, {PRIMAL, PRIMAL, PRIMAL}, {PRIMAL, PRIMAL, PRIMAL}, /* rho Vx Vy */
    {PRIMAL, PRIMAL, PRIMAL}, {PRIMAL, PRIMAL, PRIMAL}};                          /* Vz P */

static int field_ghosts(int interp) { return interp == 1 ? 2 : 4; }    /* hybrid_options.hpp:24 */
static int particle_ghosts(int interp) { return interp == 1 ? 1 : 2; } /* hybrid_options.hpp:25 */

size_t pho_field_shape(const phb_layout* L, int qty, uint32_t shape[3])
{
    size_t n = 1;
    int const g = field_ghosts(L->interp);
    for (int d = 0; d < 3; ++d)
    {
        if (d < L->dim)
            shape[d] = L->ncells[d] + (CENTERING[qty][d] == PRIMAL ? 1u : 0u) + 2u * (uint32_t)g;
        else
            shape[d] = 1;
        n *= shape[d];
    }
    return n;
}

typedef struct {
    const double* p;
    uint32_t s[3];
} fld; /* C-order view, ndarray_vector.hpp:36-51 */

static fld view(const phb_layout* L, const double* p, int qty)
{
    fld f;
    f.p = p;
    pho_field_shape(L, qty, f.s);
    return f;
}
static inline size_t at(const fld* f, int i, int j, int k)
{
    return ((size_t)i * f->s[1] + (size_t)j) * f->s[2] + (size_t)k;
}
/* physicalStart/End, ghostEnd: gridlayout.hpp:1386-1494 */
static int phys_start(const phb_layout* L) { return field_ghosts(L->interp); }
static int phys_end(const phb_layout* L, int qty, int d)
{
    return phys_start(L) + (int)L->ncells[d] - (CENTERING[qty][d] == DUAL ? 1 : 0);
}
static int ghost_end(const phb_layout* L, int qty, int d) { return phys_end(L, qty, d) + field_ghosts(L->interp); }

static int in_box(const int* c, const phb_box* b, int dim)
{
    for (int d = 0; d < dim; ++d)
        if (c[d] < b->lower[d] || c[d] > b->upper[d])
            return 0;
    return 1;
}

/* ---- B-spline start index and weights: interpolator.hpp:54-125 (Weighter), :381-406
 * (indexAndWeights_), :519-549 (computeStartLeftShift) */
int pho_weights(int order, int dual, uint32_t l, double delta, double* w)
{
    int shift;
    if (order == 1)
        shift = dual ? (delta < .5 ? 1 : 0) : 0;
    else if (order == 2)
        shift = dual ? 1 : (delta < .5 ? 1 : 0);
    else
        shift = dual ? (delta < .5 ? 2 : 1) : 1;
    uint32_t const start = l - (uint32_t)shift;
    double x             = (double)l + delta; /* iCell[iDim] + delta[iDim], iCell is uint32 */
    if (dual)
        x -= .5;
    int const si = (int)start;
    if (order == 1)
    {
        w[1] = x - (double)si;
        w[0] = 1. - w[1];
    }
    else if (order == 2)
    {
        int const index    = si + 1;
        double const d     = (double)index - x;
        double const coef1 = 0.5 + d, coef2 = d, coef3 = 0.5 - d;
        w[0] = 0.5 * coef1 * coef1;
        w[1] = 0.75 - coef2 * coef2;
        w[2] = 0.5 * coef3 * coef3;
    }
    else
    {
        double const _4_over_3 = 4. / 3., _2_over_3 = 2. / 3.;
        double const index     = (double)si - x;
        double const coef1 = 1. + 0.5 * index, coef2 = index + 1, coef3 = index + 2;
        double const coef4     = 1. - 0.5 * (index + 3);
        double const coef2_sq = coef2 * coef2, coef2_cub = coef2_sq * coef2;
        double const coef3_sq = coef3 * coef3, coef3_cub = coef3_sq * coef3;
        w[0] = _4_over_3 * coef1 * coef1 * coef1;
        w[1] = _2_over_3 - coef2_sq - 0.5 * coef2_cub;
        w[2] = _2_over_3 - coef3_sq + 0.5 * coef3_cub;
        w[3] = _4_over_3 * coef4 * coef4 * coef4;
    }
    return si;
}

typedef struct {
    int start[2][3];  /* [centering][dir] */
    double w[2][3][4];
} iw_t;

static void index_and_weights(const phb_layout* L, const int* icell, const double* delta, iw_t* iw, int both)
{
    int const g = field_ghosts(L->interp);
    for (int c = both ? 0 : PRIMAL; c <= (both ? DUAL : PRIMAL); ++c)
        for (int d = 0; d < L->dim; ++d)
        {
            /* AMRToLocal, gridlayout.hpp:746-763 */
            uint32_t const l = (uint32_t)(icell[d] - (L->amr_lower[d] - g));
            iw->start[c][d]  = pho_weights(L->interp, c == DUAL, l, delta[d], iw->w[c][d]);
        }
}

/* MeshToParticle<1|2|3>::operator(), interpolator.hpp:152-264 */
static double gather1(const phb_layout* L, const fld* f, int qty, const iw_t* iw)
{
    int const n = L->interp + 1, dim = L->dim;
    int const cx = CENTERING[qty][0], cy = CENTERING[qty][1], cz = CENTERING[qty][2];
    double F = 0.;
    if (dim == 1)
    {
        for (int ix = 0; ix < n; ++ix)
            F += f->p[at(f, 0, 0, iw->start[cx][0] + ix)] * iw->w[cx][0][ix];
        return F;
    }
    if (dim == 2)
    {
        for (int ix = 0; ix < n; ++ix)
        {
            double Y = 0.;
            for (int iy = 0; iy < n; ++iy)
                Y += f->p[at(f, 0, iw->start[cx][0] + ix, iw->start[cy][1] + iy)] * iw->w[cy][1][iy];
            F += Y * iw->w[cx][0][ix];
        }
        return F;
    }
    for (int ix = 0; ix < n; ++ix)
    {
        double Y = 0.;
        for (int iy = 0; iy < n; ++iy)
        {
            double Z = 0.;
            for (int iz = 0; iz < n; ++iz)
                Z += f->p[at(f, iw->start[cx][0] + ix, iw->start[cy][1] + iy, iw->start[cz][2] + iz)]
                     * iw->w[cz][2][iz];
            Y += Z * iw->w[cy][1][iy];
        }
        F += Y * iw->w[cx][0][ix];
    }
    return F;
}
/* 1-D/2-D arrays are stored with shape {1,1,N} / {1,N,M}: remap so that `at` works */
static fld view_nd(const phb_layout* L, const double* p, int qty)
{
    fld f = view(L, p, qty);
    if (L->dim == 1)
    {
        f.s[2] = f.s[0];
        f.s[0] = f.s[1] = 1;
    }
    else if (L->dim == 2)
    {
        f.s[2] = f.s[1];
        f.s[1] = f.s[0];
        f.s[0] = 1;
    }
    return f;
}

/* Interpolator::operator()(particle, em, layout), interpolator.hpp:420-456 */
static void gather_eb(const phb_layout* L, const fld* Ef, const fld* Bf, const int* icell, const double* delta,
                      double* e, double* b)
{
    iw_t iw;
    index_and_weights(L, icell, delta, &iw, 1);
    for (int c = 0; c < 3; ++c)
        e[c] = gather1(L, &Ef[c], PHB_EX + c, &iw);
    for (int c = 0; c < 3; ++c)
        b[c] = gather1(L, &Bf[c], PHB_BX + c, &iw);
}

/* BorisPusher::advancePosition_, boris.hpp:156-172 */
static int advance_position(int dim, const double* h, const int* icell_in, const double* delta_in,
                            const double* v, int* icell_out, double* delta_out, double* ed, double* ev)
{
    for (int d = 0; d < dim; ++d)
    {
        double const t = delta_in[d] + (h[d] * v[d]);
        if (fabs(t) > 2)
        {
            *ed = t;
            *ev = v[d];
            return PHB_ERR_MOVE_TWO_CELL;
        }
        int const s  = (int)floor(t);
        delta_out[d] = t - s;
        icell_out[d] = s + icell_in[d];
    }
    return 0;
}

/* BorisPusher::accelerate_, boris.hpp:240-300 */
static void accelerate(double* v, double charge, double dto2m, const double* E, const double* B)
{
    double const coef1 = charge * dto2m;
    double velx1 = v[0] + coef1 * E[0];
    double vely1 = v[1] + coef1 * E[1];
    double velz1 = v[2] + coef1 * E[2];
    double const rx = coef1 * B[0], ry = coef1 * B[1], rz = coef1 * B[2];
    double const rx2 = rx * rx, ry2 = ry * ry, rz2 = rz * rz;
    double const rxry = rx * ry, rxrz = rx * rz, ryrz = ry * rz;
    double const invDet = 1. / (1. + rx2 + ry2 + rz2);
    double const mxx = 1. + rx2 - ry2 - rz2;
    double const mxy = 2. * (rxry + rz);
    double const mxz = 2. * (rxrz - ry);
    double const myx = 2. * (rxry - rz);
    double const myy = 1. + ry2 - rx2 - rz2;
    double const myz = 2. * (ryrz + rx);
    double const mzx = 2. * (rxrz + ry);
    double const mzy = 2. * (ryrz - rx);
    double const mzz = 1. + rz2 - rx2 - ry2;
    double const velx2 = (mxx * velx1 + mxy * vely1 + mxz * velz1) * invDet;
    double const vely2 = (myx * velx1 + myy * vely1 + myz * velz1) * invDet;
    double const velz2 = (mzx * velx1 + mzy * vely1 + mzz * velz1) * invDet;
    velx1 = velx2 + coef1 * E[0];
    vely1 = vely2 + coef1 * E[1];
    velz1 = velz2 + coef1 * E[2];
    v[0] = velx1;
    v[1] = vely1;
    v[2] = velz1;
}

/* BorisPusher::move, boris.hpp:93-138 (prePushStep_ :180-216, postPushStep_ :218-234);
 * setMeshAndTimeStep :143-148 */
int pho_push(const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B, const phb_particles* in,
             phb_particles* out, double mass, double dt, const phb_box* first, double* ed, double* ev)
{
    int const dim = L->dim;
    double h[3];
    for (int d = 0; d < dim; ++d)
        h[d] = 0.5 * dt / L->dx[d];
    double const dto2m = 0.5 * dt / mass;
    fld Ef[3], Bf[3];
    for (int c = 0; c < 3; ++c)
    {
        Ef[c] = view_nd(L, E->comp[c], PHB_EX + c);
        Bf[c] = view_nd(L, B->comp[c], PHB_BX + c);
    }
    double dd = 0, dv = 0;
    for (size_t i = 0; i < in->n; ++i)
    {
        int ic[3] = {0, 0, 0}, ic2[3];
        double de[3] = {0, 0, 0}, de2[3], v[3], e[3], b[3];
        for (int d = 0; d < dim; ++d)
        {
            ic[d] = in->icell[d][i];
            de[d] = in->delta[d][i];
        }
        for (int c = 0; c < 3; ++c)
            v[c] = in->v[c][i];
        double const charge = in->charge[i];
        out->charge[i]      = charge;
        out->weight[i]      = in->weight[i];
        int rc              = advance_position(dim, h, ic, de, v, ic2, de2, &dd, &dv);
        if (rc)
            goto fail;
        if (!first || in_box(ic2, first, dim))
        {
            gather_eb(L, Ef, Bf, ic2, de2, e, b);
            accelerate(v, charge, dto2m, e, b);
            rc = advance_position(dim, h, ic2, de2, v, ic, de, &dd, &dv);
            if (rc)
                goto fail;
        }
        else
            for (int d = 0; d < dim; ++d)
            {
                ic[d] = ic2[d];
                de[d] = de2[d];
            }
        for (int d = 0; d < dim; ++d)
        {
            out->icell[d][i] = ic[d];
            out->delta[d][i] = de[d];
        }
        for (int c = 0; c < 3; ++c)
            out->v[c][i] = v[c];
    }
    out->n = in->n;
    return 0;
fail:
    if (ed)
        *ed = dd;
    if (ev)
        *ev = dv;
    return PHB_ERR_MOVE_TWO_CELL;
}

int pho_gather(const phb_layout* L, const phb_vecfield* E, const phb_vecfield* B, const phb_particles* in,
               double* eb)
{
    fld Ef[3], Bf[3];
    for (int c = 0; c < 3; ++c)
    {
        Ef[c] = view_nd(L, E->comp[c], PHB_EX + c);
        Bf[c] = view_nd(L, B->comp[c], PHB_BX + c);
    }
    for (size_t i = 0; i < in->n; ++i)
    {
        int ic[3]    = {0, 0, 0};
        double de[3] = {0, 0, 0};
        for (int d = 0; d < L->dim; ++d)
        {
            ic[d] = in->icell[d][i];
            de[d] = in->delta[d][i];
        }
        gather_eb(L, Ef, Bf, ic, de, eb + 6 * i, eb + 6 * i + 3);
    }
    return 0;
}

/* Interpolator::operator()(range, particleDensity, chargeDensity, flux, layout, coef),
 * interpolator.hpp:468-504 with ParticleToMesh<dim> :278-363 */
int pho_deposit(const phb_layout* L, const phb_particles* P, size_t first, size_t last, double* rho_n,
                double* rho_q, const phb_vecfield* flux, double coef, const phb_box* sel, int nsel)
{
    int const dim = L->dim, n = L->interp + 1;
    fld f   = view_nd(L, rho_n, PHB_RHO);
    double* dst[5] = {rho_n, rho_q, flux->comp[0], flux->comp[1], flux->comp[2]};
    for (size_t i = first; i < last; ++i)
    {
        int ic[3]    = {0, 0, 0};
        double de[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d)
        {
            ic[d] = P->icell[d][i];
            de[d] = P->delta[d][i];
        }
        if (nsel > 0)
        {
            int ok = 0;
            for (int b = 0; b < nsel && !ok; ++b)
                ok = in_box(ic, &sel[b], dim);
            if (!ok)
                continue;
        }
        iw_t iw;
        index_and_weights(L, ic, de, &iw, 0);
        double const q[5] = {1., P->charge[i], P->v[0][i], P->v[1][i], P->v[2][i]};
        for (int k = 0; k < 5; ++k)
        {
            double const deposit = q[k] * P->weight[i] * coef;
            if (dim == 1)
                for (int ix = 0; ix < n; ++ix)
                    dst[k][at(&f, 0, 0, iw.start[PRIMAL][0] + ix)] += deposit * iw.w[PRIMAL][0][ix];
            else if (dim == 2)
                for (int ix = 0; ix < n; ++ix)
                    for (int iy = 0; iy < n; ++iy)
                        dst[k][at(&f, 0, iw.start[PRIMAL][0] + ix, iw.start[PRIMAL][1] + iy)]
                            += deposit * iw.w[PRIMAL][0][ix] * iw.w[PRIMAL][1][iy];
            else
                for (int ix = 0; ix < n; ++ix)
                    for (int iy = 0; iy < n; ++iy)
                        for (int iz = 0; iz < n; ++iz)
                            dst[k][at(&f, iw.start[PRIMAL][0] + ix, iw.start[PRIMAL][1] + iy,
                                      iw.start[PRIMAL][2] + iz)]
                                += deposit * iw.w[PRIMAL][0][ix] * iw.w[PRIMAL][1][iy] * iw.w[PRIMAL][2][iz];
        }
    }
    return 0;
}

/* ---- binning: our canonical cell order replacing CellMap bookkeeping (cellmap.hpp:411-470) and
 * the selectors of UpdaterSelectionBoxing (ion_updater.hpp:119-163). Key definition: see
 * phb_bin in include/phare_b200.h. Stable counting sort. */
static size_t box_volume(const phb_box* b, int dim)
{
    size_t v = 1;
    for (int d = 0; d < dim; ++d)
        v *= (size_t)(b->upper[d] - b->lower[d] + 1);
    return v;
}
static size_t rowmajor(const int* c, const phb_box* b, int dim)
{
    size_t k = 0;
    for (int d = 0; d < dim; ++d)
        k = k * (size_t)(b->upper[d] - b->lower[d] + 1) + (size_t)(c[d] - b->lower[d]);
    return k;
}
static phb_box grow(const phb_box* b, int dim, int w)
{
    phb_box g = *b;
    for (int d = 0; d < dim; ++d)
    {
        g.lower[d] -= w;
        g.upper[d] += w;
    }
    return g;
}
size_t pho_bin_nkeys(const phb_layout* L, const phb_box* domain)
{
    phb_box const G = grow(domain, L->dim, particle_ghosts(L->interp));
    return box_volume(domain, L->dim) + box_volume(&G, L->dim) + 1;
}
static size_t bin_key(const phb_layout* L, const int* c, const phb_box* domain, const phb_box* G,
                      const phb_box* keep, int nkeep, size_t Nd, size_t Ng)
{
    if (in_box(c, domain, L->dim))
        return rowmajor(c, domain, L->dim);
    if (in_box(c, G, L->dim))
        for (int b = 0; b < nkeep; ++b)
            if (in_box(c, &keep[b], L->dim))
                return Nd + rowmajor(c, G, L->dim);
    return Nd + Ng;
}
static void copy_particle(int dim, const phb_particles* s, size_t i, phb_particles* d, size_t j)
{
    for (int k = 0; k < dim; ++k)
    {
        d->icell[k][j] = s->icell[k][i];
        d->delta[k][j] = s->delta[k][i];
    }
    for (int k = 0; k < 3; ++k)
        d->v[k][j] = s->v[k][i];
    d->weight[j] = s->weight[i];
    d->charge[j] = s->charge[i];
}
int pho_bin(const phb_layout* L, const phb_particles* in, phb_particles* out, const phb_box* domain,
            const phb_box* keep, int nkeep, uint32_t* cell_start, size_t counts[3])
{
    int const dim   = L->dim;
    phb_box const G = grow(domain, dim, particle_ghosts(L->interp));
    size_t const Nd = box_volume(domain, dim), Ng = box_volume(&G, dim), nk = Nd + Ng + 1;
    if (out->capacity < in->n)
        return PHB_ERR_CAPACITY;
    memset(cell_start, 0, (nk + 1) * sizeof(uint32_t));
    size_t* key = (size_t*)malloc((in->n ? in->n : 1) * sizeof(size_t));
    for (size_t i = 0; i < in->n; ++i)
    {
        int c[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d)
            c[d] = in->icell[d][i];
        key[i] = bin_key(L, c, domain, &G, keep, nkeep, Nd, Ng);
        cell_start[key[i] + 1]++;
    }
    for (size_t k = 0; k < nk; ++k)
        cell_start[k + 1] += cell_start[k];
    uint32_t* cur = (uint32_t*)malloc(nk * sizeof(uint32_t));
    memcpy(cur, cell_start, nk * sizeof(uint32_t));
    for (size_t i = 0; i < in->n; ++i)
        copy_particle(dim, in, i, out, cur[key[i]]++);
    counts[0] = cell_start[Nd];
    counts[1] = cell_start[Nd + Ng] - cell_start[Nd];
    counts[2] = cell_start[Nd + Ng + 1] - cell_start[Nd + Ng];
    out->n    = counts[0] + counts[1];
    free(cur);
    free(key);
    return 0;
}

/* ParticleArray::export_particles (particle_array.hpp:135-160) / ParticlesData pack with periodic
 * shift (particles_data.hpp:702-784) */
int pho_export(const phb_layout* L, const phb_particles* src, size_t first, size_t last, const phb_box* box,
               const phb_box* minus, const int shift[3], phb_particles* dst, size_t* appended)
{
    int const dim = L->dim;
    size_t n      = 0;
    for (size_t i = first; i < last; ++i)
    {
        int c[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d)
            c[d] = src->icell[d][i];
        if (!in_box(c, box, dim) || (minus && in_box(c, minus, dim)))
            continue;
        if (dst->n >= dst->capacity)
            return PHB_ERR_CAPACITY;
        copy_particle(dim, src, i, dst, dst->n);
        if (shift)
            for (int d = 0; d < dim; ++d)
                dst->icell[d][dst->n] += shift[d];
        dst->n++;
        n++;
    }
    if (appended)
        *appended = n;
    return 0;
}

/* ---- stencils. deriv: gridlayout.hpp:550-630 with next/prev tables :1319-1337 */
static double deriv(const phb_layout* L, const fld* f, int qty, int dir, int i, int j, int k)
{
    int idx[3] = {i, j, k};
    /* map logical direction onto the (1,1,N)/(1,N,M)/(N,M,K) storage of view_nd */
    int const a   = dir + (3 - L->dim);
    int nx[3]     = {idx[0], idx[1], idx[2]}, pv[3] = {idx[0], idx[1], idx[2]};
    if (CENTERING[qty][dir] == PRIMAL)
        nx[a] += 1; /* nextPrimal_ = 1, prevPrimal_ = 0 */
    else
        pv[a] -= 1; /* nextDual_ = 0, prevDual_ = -1 */
    double const next = f->p[at(f, nx[0], nx[1], nx[2])];
    double const prev = f->p[at(f, pv[0], pv[1], pv[2])];
    return (1. / L->dx[dir]) * (next - prev);
}
/* laplacian: gridlayout.hpp:638-700 */
static double laplacian(const phb_layout* L, const fld* f, int i, int j, int k)
{
    double lap = 0;
    for (int dir = 0; dir < L->dim; ++dir)
    {
        int const a = dir + (3 - L->dim);
        int p[3] = {i, j, k}, n[3] = {i, j, k};
        p[a] -= 1;
        n[a] += 1;
        double const prev = f->p[at(f, p[0], p[1], p[2])];
        double const here = f->p[at(f, i, j, k)];
        double const next = f->p[at(f, n[0], n[1], n[2])];
        double const inv  = 1. / L->dx[dir];
        double const l    = inv * inv * (next - 2.0 * here + prev);
        lap               = dir == 0 ? l : lap + l;
    }
    return lap;
}
/* project<stencil>: gridlayout.hpp:784-796 ; stencils gridlayout_hybrid_yee.hpp:350-749.
 * kind per direction: 0 none, 1 PrimalToDual {0,+1}, 2 DualToPrimal {-1,0}; directions >= dim
 * degenerate to the single point with coefficient 1 (directionalInterp :350-356) */
static double project(const phb_layout* L, const fld* f, int kx, int ky, int kz, int i, int j, int k)
{
    int kind[3] = {kx, ky, kz};
    int off[3][2], cnt[3];
    double cf[3];
    for (int d = 0; d < 3; ++d)
    {
        if (d < L->dim && kind[d])
        {
            int const base = kind[d] == 1;
            off[d][0]      = base - 1;
            off[d][1]      = base;
            cnt[d]         = 2;
            cf[d]          = 0.5;
        }
        else
        {
            off[d][0] = 0;
            cnt[d]    = 1;
            cf[d]     = 1.0;
        }
    }
    int const sh   = 3 - L->dim;
    double result  = 0.;
    int const id[3] = {i, j, k};
    for (int a = 0; a < cnt[0]; ++a)
        for (int b = 0; b < cnt[1]; ++b)
            for (int c = 0; c < cnt[2]; ++c)
            {
                int p[3]       = {id[0], id[1], id[2]};
                int const o[3] = {off[0][a], off[1][b], off[2][c]};
                for (int d = 0; d < L->dim; ++d)
                    p[d + sh] += o[d];
                result += (cf[0] * cf[1] * cf[2]) * f->p[at(f, p[0], p[1], p[2])];
            }
    return result;
}

/* loops of evalOnBox_ (gridlayout.hpp:1253-1284) over storage-mapped indices */
#define FOR_BOX(L, lo, hi)                                                                               \
    for (int i = (L->dim == 3 ? lo[0] : 0); i <= (L->dim == 3 ? hi[0] : 0); ++i)                         \
        for (int j = (L->dim >= 2 ? lo[L->dim - 2] : 0); j <= (L->dim >= 2 ? hi[L->dim - 2] : 0); ++j)   \
            for (int k = lo[L->dim - 1]; k <= hi[L->dim - 1]; ++k)

static void phys_box(const phb_layout* L, int qty, int* lo, int* hi)
{
    for (int d = 0; d < L->dim; ++d)
    {
        lo[d] = phys_start(L);
        hi[d] = phys_end(L, qty, d);
    }
}

/* Faraday::operator(), faraday.hpp:28-97 */
int pho_faraday(const phb_layout* L, const phb_vecfield* B, const phb_vecfield* E, phb_vecfield* Bnew,
                double dt)
{
    int const dim = L->dim;
    fld Ex = view_nd(L, E->comp[0], PHB_EX), Ey = view_nd(L, E->comp[1], PHB_EY),
        Ez = view_nd(L, E->comp[2], PHB_EZ);
    for (int c = 0; c < 3; ++c)
    {
        int lo[3], hi[3];
        fld b = view_nd(L, B->comp[c], PHB_BX + c);
        phys_box(L, PHB_BX + c, lo, hi);
        double* o = Bnew->comp[c];
        FOR_BOX(L, lo, hi)
        {
            size_t const p = at(&b, i, j, k);
            if (c == 0)
            {
                if (dim == 1)
                    o[p] = b.p[p];
                else if (dim == 2)
                    o[p] = b.p[p] - dt * deriv(L, &Ez, PHB_EZ, 1, i, j, k);
                else
                    o[p] = b.p[p] - dt * deriv(L, &Ez, PHB_EZ, 1, i, j, k) + dt * deriv(L, &Ey, PHB_EY, 2, i, j, k);
            }
            else if (c == 1)
            {
                if (dim < 3)
                    o[p] = b.p[p] + dt * deriv(L, &Ez, PHB_EZ, 0, i, j, k);
                else
                    o[p] = b.p[p] - dt * deriv(L, &Ex, PHB_EX, 2, i, j, k) + dt * deriv(L, &Ez, PHB_EZ, 0, i, j, k);
            }
            else
            {
                if (dim == 1)
                    o[p] = b.p[p] - dt * deriv(L, &Ey, PHB_EY, 0, i, j, k);
                else
                    o[p] = b.p[p] - dt * deriv(L, &Ey, PHB_EY, 0, i, j, k) + dt * deriv(L, &Ex, PHB_EX, 1, i, j, k);
            }
        }
    }
    return 0;
}

/* Ampere::operator(), ampere.hpp:26-95 on the ghost box shrunk by 1
 * (evalOnShrinkedGhostBox, gridlayout.hpp:1220-1231) */
int pho_ampere(const phb_layout* L, const phb_vecfield* B, phb_vecfield* J)
{
    int const dim = L->dim;
    fld Bx = view_nd(L, B->comp[0], PHB_BX), By = view_nd(L, B->comp[1], PHB_BY),
        Bz = view_nd(L, B->comp[2], PHB_BZ);
    for (int c = 0; c < 3; ++c)
    {
        int lo[3], hi[3];
        fld jf = view_nd(L, J->comp[c], PHB_JX + c);
        for (int d = 0; d < dim; ++d)
        {
            lo[d] = 0 + 1;
            hi[d] = ghost_end(L, PHB_JX + c, d) - 1;
        }
        double* o = J->comp[c];
        FOR_BOX(L, lo, hi)
        {
            size_t const p = at(&jf, i, j, k);
            if (c == 0)
            {
                if (dim == 1)
                    o[p] = 0.0;
                else if (dim == 2)
                    o[p] = deriv(L, &Bz, PHB_BZ, 1, i, j, k);
                else
                    o[p] = deriv(L, &Bz, PHB_BZ, 1, i, j, k) - deriv(L, &By, PHB_BY, 2, i, j, k);
            }
            else if (c == 1)
            {
                if (dim < 3)
                    o[p] = -deriv(L, &Bz, PHB_BZ, 0, i, j, k);
                else
                    o[p] = deriv(L, &Bx, PHB_BX, 2, i, j, k) - deriv(L, &Bz, PHB_BZ, 0, i, j, k);
            }
            else
            {
                if (dim == 1)
                    o[p] = deriv(L, &By, PHB_BY, 0, i, j, k);
                else
                    o[p] = deriv(L, &By, PHB_BY, 0, i, j, k) - deriv(L, &Bx, PHB_BX, 1, i, j, k);
            }
        }
    }
    return 0;
}

/* projection kinds per E component, gridlayout_hybrid_yee.hpp:450-749
 * row: [momentsToE, BxToE, ByToE, BzToE] each {kx,ky,kz} */
static const int PROJ[3][4][3] = {
    /* Ex */ {{1, 0, 0}, {1, 2, 2}, {0, 0, 2}, {0, 2, 0}},
    /* Ey */ {{0, 1, 0}, {0, 0, 2}, {2, 1, 2}, {2, 0, 0}},
    /* Ez */ {{0, 0, 1}, {0, 2, 0}, {2, 0, 0}, {2, 2, 1}}};

/* Ohm::operator(), ohm.hpp:48-270 */
int pho_ohm(const phb_layout* L, const double* n, const phb_vecfield* Ve, const double* Pe,
            const phb_vecfield* B, const phb_vecfield* J, phb_vecfield* Enew, double eta, double nu,
            int hyper_mode)
{
    int const dim = L->dim;
    fld nf = view_nd(L, n, PHB_RHO), pf = view_nd(L, Pe, PHB_P);
    fld V[3], Bf[3], Jf[3];
    for (int c = 0; c < 3; ++c)
    {
        V[c]  = view_nd(L, Ve->comp[c], PHB_VX + c);
        Bf[c] = view_nd(L, B->comp[c], PHB_BX + c);
        Jf[c] = view_nd(L, J->comp[c], PHB_JX + c);
    }
    for (int c = 0; c < 3; ++c)
    {
        int lo[3], hi[3];
        fld ef = view_nd(L, Enew->comp[c], PHB_EX + c);
        phys_box(L, PHB_EX + c, lo, hi);
        double* o       = Enew->comp[c];
        const int* pm   = PROJ[c][0];
        int const c1 = (c + 1) % 3, c2 = (c + 2) % 3; /* Ex: (y,z) ; Ey: (z,x) ; Ez: (x,y) */
        FOR_BOX(L, lo, hi)
        {
            /* ideal_: Ex = -vy*bz + vz*by ; Ey = -vz*bx + vx*bz ; Ez = -vx*by + vy*bx  (ohm.hpp:92-143) */
            double const v1 = project(L, &V[c1], pm[0], pm[1], pm[2], i, j, k);
            double const v2 = project(L, &V[c2], pm[0], pm[1], pm[2], i, j, k);
            double const b1 = project(L, &Bf[c1], PROJ[c][1 + c1][0], PROJ[c][1 + c1][1], PROJ[c][1 + c1][2], i, j, k);
            double const b2 = project(L, &Bf[c2], PROJ[c][1 + c2][0], PROJ[c][1 + c2][1], PROJ[c][1 + c2][2], i, j, k);
            double const ideal = -v1 * b2 + v2 * b1;
            /* pressure_ (ohm.hpp:145-187) */
            double pressure;
            if (c < dim)
            {
                double const nOnE  = project(L, &nf, pm[0], pm[1], pm[2], i, j, k);
                double const gradP = deriv(L, &pf, PHB_P, c, i, j, k);
                pressure           = -gradP / nOnE;
            }
            else
                pressure = 0.;
            /* resistive_ (ohm.hpp:189-209): JxToEx is the identity stencil, project = 0. + 1.0*J */
            double const jOnE      = 0. + 1.0 * Jf[c].p[at(&Jf[c], i, j, k)];
            double const resistive = eta * jOnE;
            /* hyperresistive_ (ohm.hpp:211-268) */
            double hyper;
            if (hyper_mode == 0)
                hyper = -nu * laplacian(L, &Jf[c], i, j, k);
            else
            {
                double const lvlCoeff = 1. / pow(4, L->level);
                double const bx = project(L, &Bf[0], PROJ[c][1][0], PROJ[c][1][1], PROJ[c][1][2], i, j, k);
                double const by = project(L, &Bf[1], PROJ[c][2][0], PROJ[c][2][1], PROJ[c][2][2], i, j, k);
                double const bz = project(L, &Bf[2], PROJ[c][3][0], PROJ[c][3][1], PROJ[c][3][2], i, j, k);
                double const nOnE = project(L, &nf, pm[0], pm[1], pm[2], i, j, k);
                double const b    = sqrt(bx * bx + by * by + bz * bz);
                hyper = -nu * (b / (nOnE + 0.1) + 1) * lvlCoeff * laplacian(L, &Jf[c], i, j, k);
            }
            o[at(&ef, i, j, k)] = ideal + pressure + resistive + hyper;
        }
    }
    return 0;
}

/* Electrons::update, electrons.hpp:102-128 (bulk velocity on the physical primal box),
 * :202-212 (Pe = Ne*Te everywhere) */
int pho_electrons_update(const phb_layout* L, const double* Ne, const phb_vecfield* Vi, const phb_vecfield* J,
                         double Te, phb_vecfield* Ve, double* Pe)
{
    fld nf = view_nd(L, Ne, PHB_RHO);
    int lo[3], hi[3];
    phys_box(L, PHB_RHO, lo, hi);
    fld Jf[3];
    for (int c = 0; c < 3; ++c)
        Jf[c] = view_nd(L, J->comp[c], PHB_JX + c);
    FOR_BOX(L, lo, hi)
    {
        size_t const p = at(&nf, i, j, k);
        for (int c = 0; c < 3; ++c)
        {
            /* J{x,y,z}ToMoments: DualToPrimal along the component's own direction */
            double const JOnV = project(L, &Jf[c], c == 0 ? 2 : 0, c == 1 ? 2 : 0, c == 2 ? 2 : 0, i, j, k);
            Ve->comp[c][p]    = Vi->comp[c][p] - JOnV / Ne[p];
        }
    }
    uint32_t s[3];
    size_t const nn = pho_field_shape(L, PHB_RHO, s);
    for (size_t p = 0; p < nn; ++p)
        Pe[p] = Ne[p] * Te;
    return 0;
}

/* Ions::computeChargeDensity / computeMassDensity / computeBulkVelocity, ions.hpp:75-145 */
int pho_ions_totals(size_t nn, int npop, const double* const* rho_n, const double* const* rho_q,
                    const phb_vecfield* flux, const double* mass, double* rho_q_tot, double* rho_m_tot,
                    phb_vecfield* V)
{
    for (size_t p = 0; p < nn; ++p)
    {
        double q = 0., m = 0., vx = 0., vy = 0., vz = 0.;
        for (int s = 0; s < npop; ++s)
        {
            q  = q + rho_q[s][p];
            m  = m + rho_n[s][p] * mass[s];
            vx = vx + flux[s].comp[0][p] * mass[s];
            vy = vy + flux[s].comp[1][p] * mass[s];
            vz = vz + flux[s].comp[2][p] * mass[s];
        }
        rho_q_tot[p]  = q;
        rho_m_tot[p]  = m;
        V->comp[0][p] = vx / m;
        V->comp[1][p] = vy / m;
        V->comp[2][p] = vz / m;
    }
    return 0;
}

/* core::average, utilities/algorithm.hpp:68-77 */
int pho_average(size_t n, const double* a, const double* b, double* avg)
{
    for (size_t i = 0; i < n; ++i)
        avg[i] = (a[i] + b[i]) * .5;
    return 0;
}

/* box copy / += / max between two arrays: FieldData::copy / packStream / unpackStream
 * (field_data.hpp:212-290), PlusEquals / SetMax operators (utilities/types.hpp:570-588) */
int pho_box_op(int dim, double* dst, const uint32_t ds[3], const uint32_t dlo[3], const double* src,
               const uint32_t ss[3], const uint32_t slo[3], const uint32_t ext[3], int op)
{
    uint32_t e[3] = {1, 1, 1}, dS[3] = {1, 1, 1}, sS[3] = {1, 1, 1}, dl[3] = {0, 0, 0}, sl[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d)
    {
        int const a = d + 3 - dim;
        e[a]  = ext[d];
        dS[a] = ds[d];
        sS[a] = ss[d];
        dl[a] = dlo[d];
        sl[a] = slo[d];
    }
    for (uint32_t i = 0; i < e[0]; ++i)
        for (uint32_t j = 0; j < e[1]; ++j)
            for (uint32_t k = 0; k < e[2]; ++k)
            {
                size_t const pd = ((size_t)(dl[0] + i) * dS[1] + (dl[1] + j)) * dS[2] + (dl[2] + k);
                size_t const ps = ((size_t)(sl[0] + i) * sS[1] + (sl[1] + j)) * sS[2] + (sl[2] + k);
                if (op == 0)
                    dst[pd] = src[ps];
                else if (op == 1)
                    dst[pd] += src[ps];
                else if (op == 3)
                    dst[pd] = NAN; /* setNaNsOnFieldGhosts */
                else
                    dst[pd] = (dst[pd] < src[ps]) ? src[ps] : dst[pd]; /* std::max(d, d0): NaN d stays, NaN d0 ignored */
            }
    return 0;
}

/* =====================================================================================================
 * coarse <-> fine level operators (refinement ratio 2)
 * ===================================================================================================== */
static inline size_t vat(const phb_field_view* v, const int idx[3])
{ /* AMRToLocal(index, box) = index - box.lower (amr_utils.hpp), C order */
    return ((size_t)(idx[0] - v->lo[0]) * v->shape[1] + (size_t)(idx[1] - v->lo[1])) * v->shape[2]
           + (size_t)(idx[2] - v->lo[2]);
}
/* toCoarseIndex, amr_utils.hpp:128-135 */
static inline int to_coarse(int i) { return (i >= 0) ? i / 2 : i / 2 + i % 2; }

/* LinearWeighter for ratio 2 (linear_weighter.cpp:9-56): primal distances {0, 1/2}; dual (even ratio):
 * {(0.5+0)*0.5, (0.5+1)*0.5} rotated by one -> {0.75, 0.25}; weights {1-d, d} */
static void linear_weights(int centering, double w[2][2])
{
    double dist[2];
    double const small = 1. / 2;
    if (centering == PRIMAL)
    {
        dist[0] = (double)0 / 2;
        dist[1] = (double)1 / 2;
    }
    else
    {
        dist[1] = (0.5 + (double)0) * small; /* after std::rotate by the middle */
        dist[0] = (0.5 + (double)1) * small;
    }
    for (int i = 0; i < 2; ++i)
    {
        w[i][0] = 1. - dist[i];
        w[i][1] = dist[i];
    }
}

/* DefaultFieldRefiner::operator() (field_refiner.hpp:58-163) with FieldRefineIndexesAndWeights
 * (field_linear_refine.hpp:55-118) */
static void refine_default(int dim, const int cen[3], const phb_field_view* C, const phb_field_view* F, const int f[3])
{
    int start[3] = {C->lo[0], C->lo[1], C->lo[2]}, iw[3] = {0, 0, 0};
    double w[3][2][2];
    for (int d = 0; d < dim; ++d)
    {
        double const shift = cen[d] == PRIMAL ? 0. : 0.5;
        start[d]           = (int)floor((double)(f[d] + shift) / 2 - shift);
        iw[d]              = abs(f[d]) % 2;
        linear_weights(cen[d], w[d]);
    }
    double value = 0.;
    if (dim == 1)
    {
        for (int sx = 0; sx < 2; ++sx)
        {
            int const c[3] = {start[0] + sx, start[1], start[2]};
            value += C->data[vat(C, c)] * w[0][iw[0]][sx];
        }
    }
    else if (dim == 2)
    {
        for (int sx = 0; sx < 2; ++sx)
        {
            double Y = 0.;
            for (int sy = 0; sy < 2; ++sy)
            {
                int const c[3] = {start[0] + sx, start[1] + sy, start[2]};
                Y += C->data[vat(C, c)] * w[1][iw[1]][sy];
            }
            value += Y * w[0][iw[0]][sx];
        }
    }
    else
    {
        for (int sx = 0; sx < 2; ++sx)
        {
            double Y = 0.;
            for (int sy = 0; sy < 2; ++sy)
            {
                double Z = 0.;
                for (int sz = 0; sz < 2; ++sz)
                {
                    int const c[3] = {start[0] + sx, start[1] + sy, start[2] + sz};
                    Z += C->data[vat(C, c)] * w[2][iw[2]][sz];
                }
                Y += Z * w[1][iw[1]][sy];
            }
            value += Y * w[0][iw[0]][sx];
        }
    }
    size_t const pf = vat(F, f);
    if (isnan(F->data[pf]))
        F->data[pf] = value;
}

/* MagneticFieldRefiner / MagneticFieldInitRefiner::operator() (magnetic_field_refiner.hpp:47-190,
 * magnetic_field_init_refiner.hpp:40-150): a fine face lying on a coarse face takes the coarse value; the primal
 * direction of the component must be even, the dual directions always qualify */
static void refine_magnetic(int dim, const int cen[3], const phb_field_view* C, const phb_field_view* F, const int f[3],
                            int only_nan)
{
    int c[3] = {C->lo[0], C->lo[1], C->lo[2]};
    for (int d = 0; d < dim; ++d)
    {
        c[d] = to_coarse(f[d]);
        if (cen[d] == PRIMAL && f[d] % 2 != 0)
            return;
    }
    size_t const pf = vat(F, f);
    if (!only_nan || isnan(F->data[pf]))
        F->data[pf] = C->data[vat(C, c)];
}

/* ElectricFieldRefiner::operator() (electric_field_refiner.hpp:42-57): 1-D :75-82 (always the coarse value),
 * 2-D :84-150, 3-D :153-338 */
static void refine_electric(int dim, const int cen[3], const phb_field_view* C, const phb_field_view* F, const int f[3])
{
    size_t const pf = vat(F, f);
    if (!isnan(F->data[pf]))
        return;
    int c[3] = {C->lo[0], C->lo[1], C->lo[2]};
    for (int d = 0; d < dim; ++d)
        c[d] = to_coarse(f[d]);
#define CV(dx, dy, dz) C->data[vat(C, (int[3]){c[0] + (dx), c[1] + (dy), c[2] + (dz)})]
    double value;
    if (dim == 1)
        value = CV(0, 0, 0);
    else if (dim == 2)
    {
        int const onX = f[0] % 2 == 0, onY = f[1] % 2 == 0;
        if (cen[0] == DUAL && cen[1] == PRIMAL) /* Ex */
            value = onY ? CV(0, 0, 0) : 0.5 * (CV(0, 0, 0) + CV(0, 1, 0));
        else if (cen[0] == PRIMAL && cen[1] == DUAL) /* Ey */
            value = onX ? CV(0, 0, 0) : 0.5 * (CV(0, 0, 0) + CV(1, 0, 0));
        else if (cen[0] == PRIMAL && cen[1] == PRIMAL) /* Ez */
        {
            if (onX && onY)
                value = CV(0, 0, 0);
            else if (onX)
                value = 0.5 * (CV(0, 0, 0) + CV(0, 1, 0));
            else if (onY)
                value = 0.5 * (CV(0, 0, 0) + CV(1, 0, 0));
            else
                value = 0.25 * (CV(0, 0, 0) + CV(1, 0, 0) + CV(0, 1, 0) + CV(1, 1, 0));
        }
        else
            return;
    }
    else
    {
        int const onX = f[0] % 2 == 0, onY = f[1] % 2 == 0, onZ = f[2] % 2 == 0;
        if (cen[0] == DUAL && cen[1] == PRIMAL && cen[2] == PRIMAL) /* Ex :166-196 */
        {
            if (onY && onZ)
                value = CV(0, 0, 0);
            else if (onY)
                value = 0.5 * (CV(0, 0, 0) + CV(0, 0, 1));
            else if (onZ)
                value = 0.5 * (CV(0, 0, 0) + CV(0, 1, 0));
            else
                value = 0.25 * (CV(0, 0, 0) + CV(0, 1, 0)) + 0.25 * (CV(0, 0, 1) + CV(0, 1, 1));
        }
        else if (cen[0] == PRIMAL && cen[1] == DUAL && cen[2] == PRIMAL) /* Ey :199-231 */
        {
            if (onX && onZ)
                value = CV(0, 0, 0);
            else if (onX)
                value = 0.5 * (CV(0, 0, 0) + CV(0, 0, 1));
            else if (onZ)
                value = 0.5 * (CV(0, 0, 0) + CV(1, 0, 0));
            else
                value = 0.25 * (CV(0, 0, 0) + CV(1, 0, 0) + CV(0, 0, 1) + CV(1, 0, 1));
        }
        else if (cen[0] == PRIMAL && cen[1] == PRIMAL && cen[2] == DUAL) /* Ez :234-334 */
        {
            if (onX && onY)
                value = CV(0, 0, 0);
            else if (onX)
                value = 0.5 * (CV(0, 0, 0) + CV(0, 1, 0));
            else if (onY)
                value = 0.5 * (CV(0, 0, 0) + CV(1, 0, 0));
            else
                value = 0.25 * (CV(0, 0, 0) + CV(1, 0, 0) + CV(0, 1, 0) + CV(1, 1, 0));
        }
        else
            return;
    }
#undef CV
    F->data[pf] = value;
}

int pho_field_refine(int dim, int op, int qty, const phb_field_view* coarse, const phb_field_view* fine,
                     const phb_box* box)
{
    if (dim < 1 || dim > 3 || qty < 0 || qty >= PHB_NQTY)
        return PHB_ERR_INVALID;
    int lo[3] = {fine->lo[0], fine->lo[1], fine->lo[2]}, hi[3] = {fine->lo[0], fine->lo[1], fine->lo[2]};
    for (int d = 0; d < dim; ++d)
    {
        lo[d] = box->lower[d];
        hi[d] = box->upper[d];
    }
    for (int i = lo[0]; i <= hi[0]; ++i)
        for (int j = lo[1]; j <= hi[1]; ++j)
            for (int k = lo[2]; k <= hi[2]; ++k)
            {
                int const f[3] = {i, j, k};
                if (op == PHB_REFINE_DEFAULT)
                    refine_default(dim, CENTERING[qty], coarse, fine, f);
                else if (op == PHB_REFINE_MAGNETIC || op == PHB_REFINE_MAGNETIC_INIT)
                    refine_magnetic(dim, CENTERING[qty], coarse, fine, f, op == PHB_REFINE_MAGNETIC);
                else if (op == PHB_REFINE_ELECTRIC)
                    refine_electric(dim, CENTERING[qty], coarse, fine, f);
                else
                    return PHB_ERR_INVALID;
            }
    return 0;
}

/* MagneticRefinePatchStrategy::postprocessRefine (magnetic_refine_patch_strategy.hpp:66-125); p_plus/p_minus/
 * d_plus/d_minus :374-377 */
static inline int p_plus(int i, int o) { return i + 2 - o; }
static inline int p_minus(int i, int o) { return i - o; }
static inline int d_plus(int i, int o) { return i + 1 - o; }
static inline int d_minus(int i, int o) { return i - o; }


/* 3-D new fine faces (magnetic_refine_patch_strategy.hpp:192-372).  Every sum of the reference is a list of eight
 * signed samples of another component at (minus|plus) positions per direction, evaluated left to right; TR_A/B/C are
 * the second-order lists, TR_STD the third-order one (identical for the six places it appears in). */
typedef struct { int sign, sx, sy, sz; } tr_term;
static const tr_term TR_A[8]   = {{+1,0,0,0},{-1,0,1,0},{-1,1,0,0},{+1,1,1,0},{+1,0,0,1},{-1,0,1,1},{-1,1,0,1},{+1,1,1,1}};
static const tr_term TR_B[8]   = {{+1,0,0,0},{+1,0,1,0},{-1,1,0,0},{-1,1,1,0},{-1,0,0,1},{-1,0,1,1},{+1,1,0,1},{+1,1,1,1}};
static const tr_term TR_C[8]   = {{+1,0,0,0},{-1,0,1,0},{+1,1,0,0},{-1,1,1,0},{-1,0,0,1},{+1,0,1,1},{-1,1,0,1},{+1,1,1,1}};
static const tr_term TR_STD[8] = {{+1,1,1,1},{-1,0,1,1},{-1,1,0,1},{-1,1,1,0},{+1,1,0,0},{+1,0,1,0},{+1,0,0,1},{-1,0,0,0}};

/* sample of component `comp` (primal along its own direction, dual along the others) */
static double tr_sample(const fld* f, const double* p, int comp, const int i[3], const int o[3], const tr_term* t)
{
    int const s[3] = {t->sx, t->sy, t->sz};
    int idx[3];
    for (int a = 0; a < 3; ++a)
    {
        if (a == comp)
            idx[a] = s[a] ? p_plus(i[a], o[a]) : p_minus(i[a], o[a]);
        else
            idx[a] = s[a] ? d_plus(i[a], o[a]) : d_minus(i[a], o[a]);
    }
    return p[at(f, idx[0], idx[1], idx[2])];
}
static double tr_sum(const fld* f, const double* p, int comp, const int i[3], const int o[3], const tr_term* list)
{
    double acc = tr_sample(f, p, comp, i, o, &list[0]); /* every list starts with a + term */
    for (int k = 1; k < 8; ++k)
    {
        double const v = tr_sample(f, p, comp, i, o, &list[k]);
        acc            = list[k].sign > 0 ? acc + v : acc - v;
    }
    return acc;
}

static int cell_excluded(int dim, const int c[3], const phb_box* ex, int nex)
{
    for (int b = 0; b < nex; ++b)
    {
        int in = 1;
        for (int d = 0; d < dim; ++d)
            in = in && c[d] >= ex[b].lower[d] && c[d] <= ex[b].upper[d];
        if (in)
            return 1;
    }
    return 0;
}

static int postprocess_3d(const phb_layout* L, const phb_vecfield* B, const phb_box* cells, const phb_box* ex, int nex)
{
    int const g = field_ghosts(L->interp);
    static const int ijk_factor[2] = {-1, 1};
    fld f[3];
    for (int c = 0; c < 3; ++c)
        f[c] = view(L, B->comp[c], PHB_BX + c);
    double const Dx = L->dx[0], Dy = L->dx[1], Dz = L->dx[2];
    for (int comp = 0; comp < 3; ++comp)
    {
        int lo[3], hi[3];
        for (int d = 0; d < 3; ++d)
        {
            lo[d] = cells->lower[d];
            hi[d] = cells->upper[d] + (CENTERING[PHB_BX + comp][d] == PRIMAL ? 1 : 0);
        }
        /* the two other components in the order the reference adds their terms, their second-order lists, and the
         * (offset index, mesh size of the numerator, the two mesh sizes of the denominator) of the third-order terms */
        int o1, o2, k4, k5;          /* components of T2/T3 ; offset directions of T4/T5 */
        const tr_term *l2, *l3;
        double num4, den4, num5, den5;
        if (comp == 0)      { o1 = 1; o2 = 2; l2 = TR_A; l3 = TR_B; k4 = 2; num4 = Dz; den4 = Dx * Dx + Dz * Dz; k5 = 1; num5 = Dy; den5 = Dx * Dx + Dy * Dy; }
        else if (comp == 1) { o1 = 0; o2 = 2; l2 = TR_A; l3 = TR_C; k4 = 0; num4 = Dx; den4 = Dx * Dx + Dy * Dy; k5 = 2; num5 = Dz; den5 = Dy * Dy + Dz * Dz; }
        else                { o1 = 0; o2 = 1; l2 = TR_B; l3 = TR_C; k4 = 1; num4 = Dy; den4 = Dy * Dy + Dz * Dz; k5 = 0; num5 = Dx; den5 = Dx * Dx + Dz * Dz; }
        /* T4 samples the component T2 does for Bx (by) but the other one for By (bz) and Bz (bx): :243-262, :303-322, :352-371 */
        int const c4 = comp == 0 ? o1 : (comp == 1 ? o2 : o1), c5 = comp == 0 ? o2 : (comp == 1 ? o1 : o2);
        for (int i = lo[0]; i <= hi[0]; ++i)
            for (int j = lo[1]; j <= hi[1]; ++j)
                for (int k = lo[2]; k <= hi[2]; ++k)
                {
                    int const amr[3] = {i, j, k};
                    if (amr[comp] % 2 == 0 || cell_excluded(3, amr, ex, nex))
                        continue;
                    int loc[3], off[3];
                    for (int d = 0; d < 3; ++d)
                    {
                        loc[d] = amr[d] - (L->amr_lower[d] - g);
                        off[d] = d == comp ? 1 : ((amr[d] % 2 == 0) ? 0 : 1);
                    }
                    int m[3] = {loc[0], loc[1], loc[2]}, q[3] = {loc[0], loc[1], loc[2]};
                    m[comp] -= 1;
                    q[comp] += 1;
                    double* const self = B->comp[comp];
                    double const t1 = 0.5 * (self[at(&f[comp], m[0], m[1], m[2])] + self[at(&f[comp], q[0], q[1], q[2])]);
                    double const t2 = 0.125 * tr_sum(&f[o1], B->comp[o1], o1, loc, off, l2);
                    double const t3 = 0.125 * tr_sum(&f[o2], B->comp[o2], o2, loc, off, l3);
                    double const t4 = (0.125 * ijk_factor[off[k4]] * num4 * num4 / den4) * tr_sum(&f[c4], B->comp[c4], c4, loc, off, TR_STD);
                    double const t5 = (0.125 * ijk_factor[off[k5]] * num5 * num5 / den5) * tr_sum(&f[c5], B->comp[c5], c5, loc, off, TR_STD);
                    self[at(&f[comp], loc[0], loc[1], loc[2])] = t1 + t2 + t3 + t4 + t5;
                }
    }
    return 0;
}

int pho_magnetic_postprocess(const phb_layout* L, const phb_vecfield* B, const phb_box* cells, const phb_box* ex, int nex)
{
    int const dim = L->dim, g = field_ghosts(L->interp);
    if (dim == 3)
        return postprocess_3d(L, B, cells, ex, nex);
    fld const bx = view(L, B->comp[0], PHB_BX), by = view(L, B->comp[1], PHB_BY);
    double *X = B->comp[0], *Y = B->comp[1];
    for (int comp = 0; comp < dim; ++comp)
    {
        /* toFieldBox: primal directions gain one node on the upper side (field_geometry.hpp:139-181) */
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d)
        {
            lo[d] = cells->lower[d];
            hi[d] = cells->upper[d] + (CENTERING[PHB_BX + comp][d] == PRIMAL ? 1 : 0);
        }
        for (int i = lo[0]; i <= hi[0]; ++i)
            for (int j = lo[1]; j <= hi[1]; ++j)
            {
                int const idx[3] = {i, j, 0};
                if (idx[comp] % 2 == 0) /* isNewFineFace :127-132 */
                    continue;
                if (cell_excluded(dim, idx, ex, nex))
                    continue;
                int const ix = i - (L->amr_lower[0] - g);                  /* GridLayout::AMRToLocal */
                int const iy = dim > 1 ? j - (L->amr_lower[1] - g) : 0;
                if (dim == 1) /* postprocessBx1d :134-141 */
                    X[at(&bx, ix, 0, 0)] = 0.5 * (X[at(&bx, ix - 1, 0, 0)] + X[at(&bx, ix + 1, 0, 0)]);
                else if (comp == 0) /* postprocessBx2d :143-166 */
                {
                    int const xo = 1, yo = (j % 2 == 0) ? 0 : 1;
                    X[at(&bx, ix, iy, 0)]
                        = 0.5 * (X[at(&bx, ix - 1, iy, 0)] + X[at(&bx, ix + 1, iy, 0)])
                          + 0.25
                                * (Y[at(&by, d_minus(ix, xo), p_minus(iy, yo), 0)]
                                   - Y[at(&by, d_minus(ix, xo), p_plus(iy, yo), 0)]
                                   - Y[at(&by, d_plus(ix, xo), p_minus(iy, yo), 0)]
                                   + Y[at(&by, d_plus(ix, xo), p_plus(iy, yo), 0)]);
                }
                else /* postprocessBy2d :168-190 */
                {
                    int const xo = (i % 2 == 0) ? 0 : 1, yo = 1;
                    Y[at(&by, ix, iy, 0)]
                        = 0.5 * (Y[at(&by, ix, iy - 1, 0)] + Y[at(&by, ix, iy + 1, 0)])
                          + 0.25
                                * (X[at(&bx, p_minus(ix, xo), d_minus(iy, yo), 0)]
                                   - X[at(&bx, p_plus(ix, xo), d_minus(iy, yo), 0)]
                                   - X[at(&bx, p_minus(ix, xo), d_plus(iy, yo), 0)]
                                   + X[at(&bx, p_plus(ix, xo), d_plus(iy, yo), 0)]);
                }
            }
    }
    return 0;
}

/* ElectricFieldCoarsener / MomentsCoarsener::operator() (electric_field_coarsener.hpp:52-146,
 * moments_coarsener.hpp:44-80); fineStartIndex = coarseIndex * 2 */
int pho_field_coarsen(int dim, int op, int qty, const phb_field_view* fine, const phb_field_view* coarse,
                      const phb_box* box)
{
    if (dim < 1 || dim > 3 || qty < 0 || qty >= PHB_NQTY)
        return PHB_ERR_INVALID;
    const int* cen = CENTERING[qty];
    int lo[3] = {coarse->lo[0], coarse->lo[1], coarse->lo[2]}, hi[3] = {coarse->lo[0], coarse->lo[1], coarse->lo[2]};
    for (int d = 0; d < dim; ++d)
    {
        lo[d] = box->lower[d];
        hi[d] = box->upper[d];
    }
    int dual_dir = -1, ndual = 0;
    for (int d = 0; d < dim; ++d)
        if (cen[d] == DUAL)
        {
            dual_dir = d;
            ++ndual;
        }
    if (op == PHB_COARSEN_MOMENTS && ndual)
        return PHB_ERR_INVALID; /* assert(primal) */
    if (op == PHB_COARSEN_ELECTRIC && ndual > 1)
        return PHB_ERR_INVALID; /* "no electric field should end up here" */
    for (int i = lo[0]; i <= hi[0]; ++i)
        for (int j = lo[1]; j <= hi[1]; ++j)
            for (int k = lo[2]; k <= hi[2]; ++k)
            {
                int const c[3] = {i, j, k};
                int f0[3]      = {fine->lo[0], fine->lo[1], fine->lo[2]};
                for (int d = 0; d < dim; ++d)
                    f0[d] = c[d] * 2;
                double value;
                if (op == PHB_COARSEN_MOMENTS || ndual == 0)
                    value = fine->data[vat(fine, f0)];
                else
                {
                    int f1[3] = {f0[0], f0[1], f0[2]};
                    f1[dual_dir] += 1;
                    if (dim == 1) /* :69-72 : (fine(start+1) + fine(start)) */
                        value = 0.5 * (fine->data[vat(fine, f1)] + fine->data[vat(fine, f0)]);
                    else
                        value = 0.5 * (fine->data[vat(fine, f0)] + fine->data[vat(fine, f1)]);
                }
                coarse->data[vat(coarse, c)] = value;
            }
    return 0;
}

int pho_box_fill(int dim, double* dst, const uint32_t ds[3], const uint32_t dlo[3], const uint32_t ext[3], double value)
{
    uint32_t e[3] = {1, 1, 1}, dS[3] = {1, 1, 1}, dl[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d)
    {
        int const a = d + 3 - dim;
        e[a]  = ext[d];
        dS[a] = ds[d];
        dl[a] = dlo[d];
    }
    for (uint32_t i = 0; i < e[0]; ++i)
        for (uint32_t j = 0; j < e[1]; ++j)
            for (uint32_t k = 0; k < e[2]; ++k)
                dst[((size_t)(dl[0] + i) * dS[1] + (dl[1] + j)) * dS[2] + (dl[2] + k)] = value;
    return 0;
}

/* PlusEqualsProduct (types.hpp:584-588): d += d0 * o */
int pho_axpy(size_t n, double* dst, const double* src, double coef)
{
    for (size_t i = 0; i < n; ++i)
        dst[i] += src[i] * coef;
    return 0;
}
